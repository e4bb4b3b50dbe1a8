"""BASELINE INFRASTRUCTURE ONLY — place an UN-MODIFIED copy of the reference's Moment-Query python tree (MQ/libs, MQ/configs)
under baseline/_ref/MQ so that bench.py's `--impl reference` arm and its cpu / eager-GPU baseline legs can run the reference's
OWN modules on the GPU box, where /root/reference does not exist.  baseline/_ref is git-ignored (never part of the history)
and not gpurun-ignored (it travels with the snapshot).  Nothing in vilco_b200/ imports it.

    python -m baseline.vendor_ref        (run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil

SRC = os.path.join(os.environ.get("VILCO_REFERENCE_SRC", "/root/reference"), "MQ")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "MQ")


def vendor(force=False):
    if not os.path.isdir(os.path.join(SRC, "libs")):
        return os.path.isdir(os.path.join(DST, "libs"))
    if os.path.isdir(os.path.join(DST, "libs")) and not force:
        return True
    for sub in ("libs", "configs"):
        d = os.path.join(DST, sub)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(SRC, sub), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "build", "*.so", "*.o"))
    return True


if __name__ == "__main__":
    print("vendored:", vendor(force=True), DST)
