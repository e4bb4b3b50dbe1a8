"""Per-operator backward parity on the GPU: every primitive of vilco_b200/backward.py against torch autograd in fp32
(linear / k=3 conv data + weight gradients through the MN-major GEMM operands and split-K, LayerNorm (+ReLU, +residual),
depthwise conv + LN for q/k/v at stride 1 and 2, GELU, max-pool, global / cross attention at several shapes).
The cases live in tools/bwd_probe.py (also usable stand-alone); tolerance 2e-4 relative, measured ~1e-5."""
import contextlib
import io
import os
import runpy

import pytest

pytestmark = pytest.mark.gpu
PRECISION = "fp16x3"   # the tolerances below state the exact (split-operand) arithmetic; see conftest._precision_mode


def test_backward_primitives_match_torch_autograd():
    buf = io.StringIO()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "bwd_probe.py")
    with contextlib.redirect_stdout(buf):
        try:
            runpy.run_path(path, run_name="__main__")
        except SystemExit:      # the probe ends with sys.exit(number of failures)
            pass
    out = buf.getvalue()
    bad = [l for l in out.splitlines() if l.startswith("BAD")]
    assert "bad 0" in out and not bad, "\n".join(bad[:10])
    assert out.count("OK ") >= 40

