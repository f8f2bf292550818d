#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's own hot-path modules, so that bench.py can time the REAL
reference (not only the C port) on the GPU box's host cores.

The reference is Python; /root/reference exists only in the build container.  `__graft_entry__.build()`
runs this there: the three files of the path are copied verbatim into oracle/_ref/ (git-ignored — they never
enter the history — but not gpurun-ignored, so they travel with the snapshot like the built .so files):

    UAV_Navigation/path_planning_GA/path_planner_hrmaga.py     (HR)
    UAV_Navigation/path_planning_GA/GA_complex _PF.py          (PF)
    uav-backend/app/routers/GA_enhanced.py                     (GAE)

Nothing under oracle/_ref/ is imported by the product package or by the parity tests; its only consumer is
bench.py's CPU legs (`cpu_baseline` kind "reference", `--impl reference`) through oracle/ref_py.py.
Without /root/reference (the GPU box) this is a no-op and whatever was copied earlier is used.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["UAV_Navigation/path_planning_GA/path_planner_hrmaga.py",
         "UAV_Navigation/path_planning_GA/GA_complex _PF.py",
         "uav-backend/app/routers/GA_enhanced.py"]


def make(src_root=None):
    src_root = src_root or os.environ.get("TRACEFLY_REFERENCE", "/root/reference")
    if not all(os.path.exists(os.path.join(src_root, f)) for f in FILES):
        return False
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, f), dst)
    with open(os.path.join(DST, "README"), "w") as fh:
        fh.write("verbatim copies of the reference's GA modules made by oracle/make_ref.py; git-ignored; "
                 "used only by bench.py's CPU legs\n")
    return True


if __name__ == "__main__":
    ok = make(sys.argv[1] if len(sys.argv) > 1 else None)
    print("oracle/_ref written" if ok else "reference tree not found: oracle/_ref left as it is")
