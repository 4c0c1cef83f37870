"""Tensor-core all-pairs kernel, isolation timings (dev build: python -m mdtraj_b200.build --dev --force): the whole
kernel, without the solve (B200RMSD_TC_DEBUG=0x100), operand delivery alone (0x300), for the single-CTA and the CTA-pair
geometry.   python tools/ap_isolate.py [F] [N]"""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import mdtraj_b200 as mdb
from mdtraj_b200 import _capi
libs = [a[6:] for a in sys.argv[1:] if a.startswith("--lib=")]   # variants/*.so built by tools/build_variants.sh
sys.argv = [a for a in sys.argv if not a.startswith("--lib=")]
if libs:
    _capi.LIB_PATH = os.path.abspath(libs[0])
from mdtraj_b200 import allpairs as AP
from ap_time import md_like
from ap_pair_check import timed

F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda", 0)
out = torch.empty((F, F), dtype=torch.float32, device=dev)
for name, dt in (("iid", mdb.DeviceTrajectory.synthetic_iid(F, N, 1, dev)), ("md", md_like(F, N, dev))):
    prep = AP.prepare(dt)
    for pair in (0, 1):
        os.environ["B200RMSD_TC_PAIR"] = str(pair)
        res = {"lib": os.path.basename(libs[0]) if libs else None, "data": name, "pair": pair}
        for flag, key in ((None, "full"), ("0x400", "nostore"), ("0x100", "nosolve"), ("0x500", "nosolve_nostore"), ("0x300", "delivery_only"), ("0x700", "delivery_only_nostore")):
            if flag: os.environ["B200RMSD_TC_DEBUG"] = flag
            else: os.environ.pop("B200RMSD_TC_DEBUG", None)
            res[key + "_ms"] = round(timed(lambda: AP.rows(prep, 0, F, out=out)), 3)
        os.environ.pop("B200RMSD_TC_DEBUG", None)
        print(json.dumps(res), flush=True)
