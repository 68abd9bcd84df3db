"""The runner shim (cfdbench_b200/runner.py) rebinding logic, exercised against the reference tree when it
is present (build container only; the GPU box has no /root/reference, so these tests skip there)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")

CODE = r'''
import sys
sys.path.insert(0, %r)
from cfdbench_b200 import runner
runner.install(%r, stub_missing=True)
import models.fno.fno2d as ref
from models.base_model import AutoCfdModel
from models.loss import loss_name_to_fn
import cfdbench_b200.fno2d as ours
assert ref.Fno2d is ours.Fno2d, "seam not rebound"
m = ref.Fno2d(in_chan=2, out_chan=2, n_case_params=5, loss_fn=loss_name_to_fn("nmse"), num_layers=4,
              hidden_dim=32, modes1=12, modes2=12, device="cpu")
assert isinstance(m, AutoCfdModel), "must subclass the reference AutoCfdModel (test_multistep.py:109)"
assert m.loss_fn.get_score_names() == ["mse", "rmse", "mae", "nmse"]
# the factory the scripts use picks the rebound class up (utils/autoregressive.py:10,114-125)
try:
    import utils.autoregressive as ua
    assert ua.Fno2d is ours.Fno2d
    print("factory-ok")
except Exception as e:
    print("factory-skip", type(e).__name__, e)
print("ok")
'''


def test_runner_rebinds_the_seam_and_subclasses_reference_base():
    out = subprocess.run([sys.executable, "-c", CODE % (ROOT, REF)], capture_output=True, text=True,
                         env={**os.environ, "PYTHONDONTWRITEBYTECODE": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("ok"), out.stdout
