"""Driver of tests/test_app_shim_cpu.py (run in a SUBPROCESS: the reference app uses top-level module names such as
``config``, ``utils``, ``models``, ``main`` that must not leak into the test session).

    python tests/app_shim_driver.py make-pickle <dir>     # the reference's own module, pickled the way the app ships it
    python tests/app_shim_driver.py serve <dir>           # unmodified app/src/main.py under TestClient, through the shim

Prints one JSON object on the last line.
"""
import base64
import io
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_APP = "/root/reference/app/src"
SHIM = os.path.join(ROOT, "integration", "app_shim")
SEED_I = 1234


def make_pickle(out_dir):
    """model.pth = torch.save(<whole reference module>) as app/src/im2latex.py:11 expects; vocab.json as utils.load_vocab reads."""
    import torch
    import torchvision
    sys.path.insert(0, ROOT)
    from handwritten_math_ocr_api_b200.layout import ModelConfig
    from handwritten_math_ocr_api_b200.synthetic import synth_state_dict, synth_vocab
    orig = torchvision.models.swin_t
    torchvision.models.swin_t = lambda weights=None, **k: orig(weights=None, **k)     # no network
    sys.path.insert(0, REF_APP)
    import model_swin                                                     # the REFERENCE's module (app flavour)
    assert model_swin.__file__.startswith(REF_APP)
    cfg = ModelConfig()
    model = model_swin.FormulaRecognitionModel(cfg.vocab_size).eval()
    model.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
    torch.save(model, os.path.join(out_dir, "model.pth"))
    vocab, idx2char = synth_vocab(cfg.vocab_size)
    with open(os.path.join(out_dir, "vocab.json"), "w") as f:
        json.dump({"vocab": vocab, "idx2char": {str(k): v for k, v in idx2char.items()}}, f)
    print(json.dumps({"ok": True, "keys": len(model.state_dict())}))


class FakeEngine:
    """Stands in for handwritten_math_ocr_api_b200.model_swin.FormulaRecognitionModel in the build container (no GPU):
    replays the outputs the unmodified reference produced for the golden images (tests/golden/*.npz)."""
    loaded_keys = 0
    calls = []

    def __init__(self, vocab_size, config=None, device=None, **kw):
        import numpy as np
        import torch
        from handwritten_math_ocr_api_b200.synthetic import synth_images
        self.vocab_size, self.device = vocab_size, device
        g = np.load(os.path.join(ROOT, "tests", "golden", "swin_src_golden.npz"))
        a = np.load(os.path.join(ROOT, "tests", "golden", "swin_app_golden.npz"))
        self.images = synth_images(4, SEED_I)
        self.ys = torch.from_numpy(g["greedy_ys"])
        self.lp_sums = a["logprob_sums"]
        self.eos_id, self.sos_id, self.pad_id = 2, 1, 0

    def load_state_dict(self, sd, strict=True):
        FakeEngine.loaded_keys = len(sd)

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def generate(self, images, max_len=None, beam_size=1, return_logprobs=False, encoder_out=None):
        import torch
        rows = []
        for b in range(images.shape[0]):
            hit = [i for i in range(4) if torch.equal(images[b].cpu(), self.images[i])]
            assert hit, "the app handed the engine a tensor that is not one of the golden images"
            rows.append(hit[0])
        FakeEngine.calls.append(rows)
        tok = self.ys[rows]
        steps = tok.shape[1] - 1
        logp = torch.zeros(len(rows), steps)
        for j, i in enumerate(rows):
            seq = tok[j, 1:].tolist()
            n = seq.index(self.eos_id) + 1 if self.eos_id in seq else steps
            if i < len(self.lp_sums):
                logp[j, :n] = float(self.lp_sums[i]) / n        # only the SUM enters the confidence (im2latex.py:37,55)
        return tok, steps, (logp if return_logprobs else None)


def serve(model_dir):
    import numpy as np
    sys.path.insert(0, ROOT)
    # third-party services of the app that do not exist here: Cloud Logging, Redis
    g = types.ModuleType("google"); gc = types.ModuleType("google.cloud"); gl = types.ModuleType("google.cloud.logging")
    gl.Client = lambda *a, **k: types.SimpleNamespace(setup_logging=lambda: None)
    g.cloud = gc; gc.logging = gl
    sys.modules.update({"google": g, "google.cloud": gc, "google.cloud.logging": gl})
    redis = types.ModuleType("redis")
    redis.from_url = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no redis in the test"))
    redis.Redis = object
    sys.modules["redis"] = redis
    for k in ("GOOGLE_CLOUD_PROJECT", "MODEL_API_KEY", "REDIS_URL"):
        os.environ.pop(k, None)
    os.environ["ENVIRONMENT"] = "development"

    sys.path.insert(0, REF_APP)
    sys.path.insert(0, SHIM)                      # the shim goes IN FRONT of app/src
    import config as ref_config
    assert ref_config.__file__.startswith(REF_APP)
    ref_config.Config.MODEL_DIR = __import__("pathlib").Path(model_dir)     # app/trained-model holds no checkpoint (D4)
    import handwritten_math_ocr_api_b200.model_swin as engine_mod
    engine_mod.FormulaRecognitionModel = FakeEngine                         # the build container has no GPU
    import im2latex
    import model_swin
    import preprocess
    assert im2latex.__file__.startswith(SHIM) and model_swin.__file__.startswith(SHIM) and preprocess.__file__.startswith(SHIM)
    import main                                                             # the UNMODIFIED reference app
    assert main.__file__.startswith(REF_APP)
    from fastapi.testclient import TestClient
    from PIL import Image
    from handwritten_math_ocr_api_b200.synthetic import synth_images_u8
    golden = np.load(os.path.join(ROOT, "tests", "golden", "swin_app_golden.npz"))
    u8 = synth_images_u8(2, SEED_I).numpy()

    def png(i):
        buf = io.BytesIO()
        Image.fromarray(u8[i], mode="L").save(buf, format="PNG")
        return buf.getvalue()

    out = {}
    with TestClient(main.app) as client:                                    # runs the lifespan: initialize_model()
        out["model_loaded"] = main.model is not None and isinstance(main.model, FakeEngine)
        out["state_dict_keys"] = FakeEngine.loaded_keys
        out["health"] = client.get("/health").status_code
        r = client.post("/predict", files={"file": ("formula1.png", png(1), "image/png")})
        out["predict_status"] = r.status_code
        body = r.json()
        out["predict_formula_equal"] = body.get("formula") == str(golden["formulas"][1])
        out["predict_confidence"] = body.get("confidence")
        out["reference_confidence"] = float(golden["confidences"][1])
        r = client.post("/predict/batch", json={"images": [base64.b64encode(png(i)).decode() for i in (0, 1)]})
        out["batch_status"] = r.status_code
        res = r.json().get("results", [])
        out["batch_success"] = [x.get("success") for x in res]
        out["batch_formula_equal"] = [x.get("formula") == str(golden["formulas"][i]) for i, x in enumerate(res)]
        out["batch_confidence"] = [x.get("confidence") for x in res]
        out["reference_batch_confidence"] = [float(c) for c in golden["confidences"]]
        r = client.post("/predict", files={"file": ("empty.png", b"", "image/png")})
        out["empty_upload_status"] = r.status_code
        # tensor-batched form of the /predict/batch loop: one engine call for both images
        import torch
        tensors = torch.cat([preprocess.preprocess_image(Image.open(io.BytesIO(png(i)))) for i in (0, 1)])
        n_calls = len(FakeEngine.calls)
        pb = im2latex.predict_batch(main.model, tensors, main.vocab, main.idx2char, "cpu")
        out["predict_batch_one_call"] = len(FakeEngine.calls) == n_calls + 1 and FakeEngine.calls[-1] == [0, 1]
        out["predict_batch_formula_equal"] = [f == str(golden["formulas"][i]) for i, (f, _) in enumerate(pb)]
    print(json.dumps(out))


if __name__ == "__main__":
    {"make-pickle": make_pickle, "serve": serve}[sys.argv[1]](sys.argv[2])
