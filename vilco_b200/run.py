"""Run one of the reference's own entry scripts, unmodified, on top of vilco_b200:

    cd ViLCo/MQ
    python -m vilco_b200.run train_cl.py configs/mq_vilco.yaml --output vilco_b200
    torchrun --nproc-per-node 8 -m vilco_b200.run train_cl.py configs/mq_vilco.yaml ...
    python -m vilco_b200.run --mq-root /path/to/ViLCo/MQ eval.py configs/mq_vilco.yaml ckpt/...

What it does before handing control to the script (`runpy`, `__name__ == "__main__"`, `sys.argv` = the script's own):
puts the MQ directory first on `sys.path` and makes it the working directory (the reference opens its XLNet json and its
data files relative to it, backbones.py:132), imports `libs.utils` before `libs.modeling` (the order the reference's import
cycle needs, meta_archs.py:15 <-> train_utils.py:19) and calls `vilco_b200.compat.install()`, so that every
`from libs.modeling import make_meta_arch` / `from libs.utils import batched_nms, ANETdetection, ...` in the script binds the
sm_100a implementations.  Nothing of the reference is edited or copied.
"""
import argparse
import os
import runpy
import sys


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m vilco_b200.run", description=__doc__.split("\n\n")[0])
    ap.add_argument("--mq-root", default=None, help="the ViLCo/MQ directory (default: the current directory)")
    ap.add_argument("script", help="reference entry script, e.g. train_cl.py, train.py, train_bic.py, eval.py")
    ap.add_argument("args", nargs=argparse.REMAINDER, help="arguments of the script")
    a = ap.parse_args(argv)
    root = os.path.abspath(a.mq_root or os.getcwd())
    if not os.path.isdir(os.path.join(root, "libs", "modeling")):
        raise SystemExit(f"vilco_b200.run: {root} is not a ViLCo/MQ checkout (libs/modeling missing); use --mq-root")
    script = a.script if os.path.isabs(a.script) else os.path.join(root, a.script)
    if not os.path.isfile(script):
        script = os.path.abspath(a.script)
    if not os.path.isfile(script):
        raise SystemExit(f"vilco_b200.run: script {a.script!r} not found (looked in {root} and the current directory)")
    os.chdir(root)
    if root not in sys.path:
        sys.path.insert(0, root)
    import libs.utils      # noqa: F401  first: breaks the libs.modeling <-> libs.utils import cycle
    import libs.modeling   # noqa: F401
    from . import compat
    patched = compat.install()
    if os.environ.get("RANK", "0") == "0":
        print(f"[vilco_b200.run] {len(patched)} names of libs.modeling / libs.utils rebound to vilco_b200; running {script}",
              flush=True)
    sys.argv = [script] + list(a.args)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
