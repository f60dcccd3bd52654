"""A/B of differently built join libraries (EJB_LIB) on BASELINE configs[1] (dev tool; EJB_AB_CONFIG=4 for the 10M-cell one).
  python tools/ab_lib.py build/lib_a.so build/lib_b.so ...      ("default" = the in-tree library)"""
import json
import os
import subprocess
import sys

sys.path.insert(0, ".")
PATH = "/dev/shm/ejb_ab_inp.npz"

if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    import numpy as np
    import enclone_ranger_b200 as E
    z = np.load(PATH)
    inp = E.JoinInput(**{k: z[k] for k in z.files})
    p = E.default_params(True)
    ctx = E.JoinContext()
    ctx.upload(p, inp)
    ctx.run()
    ss = [ctx.run() for _ in range(4)]
    s = min(ss, key=lambda x: x["ms_device_total"])
    keys = ("ms_device_total", "ms_pack", "ms_stage1", "ms_stage2", "ms_stage2a", "ms_forest", "ms_finalize", "stage1_candidates", "rounds", "kernel_launches")
    print(json.dumps(dict({"lib": os.environ.get("EJB_LIB", "default")}, **{k: (round(s[k], 3) if isinstance(s[k], float) else int(s[k])) for k in keys})), flush=True)
    if os.environ.get("EJB_TRACE"):
        ctx.run()
    ctx.close()
    sys.exit(0)

import numpy as np
import enclone_ranger_b200 as E
rep = E.generate(int(os.environ.get("EJB_AB_CONFIG", "1")))
inp = rep.to_join_input().to_packed().to_packed_cdr3()
rep.close()
np.savez(PATH, **inp.a)
for lib in sys.argv[1:] or ["default"]:
    env = dict(os.environ)
    if lib != "default":
        env["EJB_LIB"] = os.path.abspath(lib)
    subprocess.run([sys.executable, __file__, "--worker"], env=env, check=False)
os.unlink(PATH)
