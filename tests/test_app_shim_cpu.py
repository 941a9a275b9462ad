"""the UNMODIFIED FastAPI app of the reference (app/src/main.py), imported
under TestClient with Cloud Logging and Redis stubbed, served through integration/app_shim (module names `im2latex`,
`model_swin`, `preprocess` resolve to the shim, everything else to the reference's own files).  The build container
has no GPU, so the engine class is replaced by a fake that replays the outputs recorded from the real reference
(tests/golden); the same (formula, confidence) pair is checked on the real engine by
tests/test_parity_gpu.py::test_im2latex_predict_against_the_reference_api_golden.

Needs /root/reference (build container only): skipped elsewhere."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

DRIVER = os.path.join(ROOT, "tests", "app_shim_driver.py")

pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/app/src/main.py"),
                                reason="the reference app only exists in the build container")


def _run(cmd, arg):
    r = subprocess.run([sys.executable, DRIVER, cmd, arg], capture_output=True, text=True, timeout=600,
                       cwd=arg)          # cwd: the reference's config.py creates directories relative to it
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_reference_app_serves_through_the_shim(tmp_path):
    pytest.importorskip("fastapi")
    pytest.importorskip("httpx")
    d = str(tmp_path)
    made = _run("make-pickle", d)
    assert made["keys"] == 517                                   # the reference module's own state_dict
    out = _run("serve", d)
    assert out["model_loaded"] and out["state_dict_keys"] == 517  # un-pickled by module path, handed to the engine whole
    assert out["health"] == 200
    assert out["predict_status"] == 200 and out["predict_formula_equal"]
    assert abs(out["predict_confidence"] - out["reference_confidence"]) < 1e-6
    assert out["batch_status"] == 200 and out["batch_success"] == [True, True]
    assert out["batch_formula_equal"] == [True, True]
    for a, b in zip(out["batch_confidence"], out["reference_batch_confidence"]):
        assert abs(a - b) < 1e-6
    assert out["empty_upload_status"] == 400
    assert out["predict_batch_one_call"] and out["predict_batch_formula_equal"] == [True, True]
