// Host build of the device QCP solvers (test infrastructure; see shim/cuda_runtime.h).
static long g_slow = 0;
#define QCP_SLOW_PATH_HOOK() (++g_slow)
#include "../../mdtraj_b200/csrc/qcp.cuh"

using namespace b200;

extern "C" {
long host_qcp_slow_count(int reset) { const long v = g_slow; if (reset) g_slow = 0; return v; }
// qcp_solve on n problems: M (n,9) float64, Ga/Gb (n), n_atoms; out msd (n), rot (n,9) float32 or NULL, degenerate flags (n)
void host_qcp_solve(const double* M, const double* Ga, const double* Gb, int n_atoms, long n, double* msd, float* rot,
                    unsigned char* degen)
{
    for (long i = 0; i < n; ++i) {
        QcpInput q;
        for (int k = 0; k < 9; ++k) q.M[k] = M[9 * i + k];
        q.Ga = Ga[i]; q.Gb = Gb[i]; q.inv_n = 1.0 / (double)n_atoms;
        bool d = false;
        msd[i] = qcp_solve(q, rot ? rot + 9 * i : nullptr, &d);
        degen[i] = d;
    }
}
// the all-pairs epilogue solvers on n problems: M (n,9) float32, Ga/Gb float32; out rmsd (n) float32
void host_qcp_msd_fast(const float* M, const float* Ga, const float* Gb, int n_atoms, long n, int f32_only, float* rmsd)
{
    const float inv_n = 1.0f / (float)n_atoms;
    for (long i = 0; i < n; ++i) {
        float m[1][9], ga[1] = {Ga[i]}, gb[1] = {Gb[i]}, r[1];
        bool act[1] = {true}, ok[1];
        for (int k = 0; k < 9; ++k) m[0][k] = M[9 * i + k];
        if (f32_only) qcp_msd_f32<1>(m, ga, gb, act, inv_n, r, ok);
        else qcp_msd_fast<1>(m, ga, gb, act, inv_n, r, ok);
        rmsd[i] = ok[0] ? r[0] : qcp_rmsd_closed(m[0], ga[0], gb[0], inv_n);  // what the epilogue does
    }
}
// round 2's epilogue solver: qcp_msd_shift first; where it does not trust itself, qcp_msd_fast, then the closed form --
// the same cascade as allpairs_tc144.cu.  M must be in the natural row order (x, y, z).  stage[i]: 0 shift, 1 fast, 2 closed.
void host_qcp_msd_shift(const float* M, const float* Ga, const float* Gb, int n_atoms, long n, float* rmsd,
                        unsigned char* stage)
{
    const float inv_n = 1.0f / (float)n_atoms;
    for (long i = 0; i < n; ++i) {
        float m[1][9], ga[1] = {Ga[i]}, gb[1] = {Gb[i]}, r[1];
        bool act[1] = {true}, ok[1];
        for (int k = 0; k < 9; ++k) m[0][k] = M[9 * i + k];
        qcp_msd_shift<1>(m, ga, gb, act, (float)n_atoms, r, ok);
        stage[i] = 0;
        if (!ok[0]) {
            qcp_msd_fast<1>(m, ga, gb, act, inv_n, r, ok);
            stage[i] = 1;
            if (!ok[0]) { r[0] = qcp_rmsd_closed(m[0], ga[0], gb[0], inv_n); stage[i] = 2; }
        }
        rmsd[i] = r[0];
    }
}
}
