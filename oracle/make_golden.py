"""Generate ``tests/golden/*`` from the UNMODIFIED reference and check the oracle against it.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python -m oracle.make_golden            # both phases, each in its own process

Phase ``src``  imports ``/root/reference/src/{model_swin,inference}.py`` (one offline patch:
``torchvision.models.swin_t(weights=None)`` because the ImageNet download at
``src/model_swin.py:17`` cannot succeed offline), loads the synthetic checkpoint of
``oracle/synth.py`` into it and records: the state-dict manifest, the parameter count, encoder
features, teacher-forced logits and ``inference.predict`` greedy strings.
Phase ``app``  imports ``/root/reference/app/src/{model_swin,im2latex}.py`` (same module names as
``src/`` => separate process) and records ``im2latex.predict`` (formula, confidence).

Each phase also evaluates the oracle restatement on the same inputs and refuses to write the
fixture if they disagree beyond fp32 reassociation noise.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
SCRATCH = os.path.join(ROOT, "baseline", "_ref", "cwd", "work")   # reference config.py mkdirs <cwd>/../checkpoints

N_IMG = 4
SEED_W, SEED_I = 0, 1234


def _patch_torchvision():
    import torchvision
    orig = torchvision.models.swin_t
    torchvision.models.swin_t = lambda weights=None, **k: orig(weights=None, **k)


def _inputs():
    sys.path.insert(0, ROOT)
    from oracle.arch import ModelConfig
    from oracle.synth import synth_images, synth_state_dict, synth_vocab
    cfg = ModelConfig()
    sd = synth_state_dict(cfg, seed=SEED_W)
    imgs = synth_images(N_IMG, seed=SEED_I)
    vocab, idx2char = synth_vocab(cfg.vocab_size)
    return cfg, sd, imgs, vocab, idx2char


def phase_src():
    os.makedirs(SCRATCH, exist_ok=True)
    os.chdir(SCRATCH)
    _patch_torchvision()
    cfg, sd, imgs, vocab, idx2char = _inputs()
    from oracle import decode as odec
    from oracle import ref_model as oref
    from oracle.arch import state_dict_layout
    from oracle.synth import state_dict_checksum

    sys.path.insert(0, os.path.join(REF, "src"))
    import model_swin                                           # the reference, unmodified
    sys.modules["model"] = model_swin                           # inference.py: `from model import ...`
    import inference as ref_inference

    torch.manual_seed(0)
    model = model_swin.FormulaRecognitionModel(cfg.vocab_size).eval()
    ref_sd = model.state_dict()
    manifest = [(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in ref_sd.items()]
    mine = [(k, list(s), d) for k, s, d in state_dict_layout(cfg)]
    assert manifest == mine, "oracle/arch.py layout differs from the reference state_dict"
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params == 37_450_293, n_params                     # README.md:89, images/mlflow-swin.png
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected

    with torch.no_grad():
        feats = model.encoder(imgs)                             # [4,30,256]
        ofeats = oref.encoder_forward(imgs, sd)
        d = (feats - ofeats).abs().max().item()
        print("encoder  max|ref-oracle| =", d, " max|ref| =", feats.abs().max().item())
        assert d < 2e-4, d
        # batch independence: image 1 alone == image 1 in the batch
        d1 = (model.encoder(imgs[1:2]) - feats[1:2]).abs().max().item()
        print("encoder  batch-independence diff =", d1)

        g = torch.Generator().manual_seed(7)
        tgt = torch.randint(4, cfg.vocab_size, (2, 6), generator=g)
        tgt[:, 0] = cfg.sos
        logits = model.decoder(feats[:2], tgt)                  # [2,6,V]
        ologits = oref.decoder_forward(feats[:2], tgt, sd, cfg)
        d = (logits - ologits).abs().max().item()
        print("decoder  max|ref-oracle| =", d, " logits std =", logits.std().item())
        assert d < 5e-4, d
        fwd = model(imgs[:2], tgt)                              # src flavour: captions[:, :-1]
        ofwd = oref.model_forward(imgs[:2], tgt, sd, cfg, drop_last=True)
        assert (fwd - ofwd).abs().max().item() < 5e-4

        # the reference's batched greedy loop, unmodified
        strings = ref_inference.predict(imgs, model, vocab, idx2char, "cpu")
        ys = odec.greedy_batched(imgs, sd, cfg)
        ostrings = [odec.ids_to_string(s.tolist(), idx2char, cfg) for s in ys]
        assert strings == ostrings, (strings, ostrings)
        yc, step_logits = odec.greedy_cached(ofeats, sd, cfg, return_logits=True)
        assert torch.equal(yc, ys), "KV-cached oracle greedy differs from full-recompute greedy"
        lens = [len(s.split()) for s in strings]
        print("greedy   steps =", ys.shape[1] - 1, " string lengths =", lens)
        # predict.py:49-67 restated (cannot be imported) - consistency with the batched loop
        single = odec.greedy_single(imgs[0:1], sd, cfg)
        # top1-top2 margin at every greedy step (for near-tie classification in the GPU tests)
        top2 = step_logits.topk(2, -1).values
        margin = (top2[..., 0] - top2[..., 1])

    np.savez_compressed(
        os.path.join(GOLD, "swin_src_golden.npz"),
        images_seed=np.int64(SEED_I), weights_seed=np.int64(SEED_W),
        checksum=np.array(state_dict_checksum(sd)),
        features=feats.numpy(), tgt=tgt.numpy(), logits=logits.numpy(),
        greedy_ys=ys.numpy(), greedy_margin=margin.numpy(),
        greedy_strings=np.array(strings), single_tokens=np.array(single, dtype=np.int64),
        versions=np.array(json.dumps({"torch": torch.__version__,
                                      "torchvision": __import__("torchvision").__version__})),
    )
    with open(os.path.join(GOLD, "state_dict_manifest.json"), "w") as f:
        json.dump({"n_params": n_params, "entries": manifest}, f)
    print("wrote swin_src_golden.npz, state_dict_manifest.json")


def phase_app():
    os.makedirs(SCRATCH, exist_ok=True)
    os.chdir(SCRATCH)
    _patch_torchvision()
    cfg, sd, imgs, vocab, idx2char = _inputs()
    from oracle import decode as odec
    sys.path.insert(0, os.path.join(REF, "app", "src"))
    import model_swin                                           # app flavour
    import im2latex as ref_im2latex
    model = model_swin.FormulaRecognitionModel(cfg.vocab_size).eval()
    model.load_state_dict(sd, strict=True)
    out = []
    for i in range(2):
        formula, conf = ref_im2latex.predict(model, imgs[i:i + 1], vocab, idx2char, "cpu")
        toks, lp, oconf = odec.greedy_api(imgs[i:i + 1], sd, cfg)
        oformula = odec.clean_latex_output(odec.tokens_to_latex(toks, idx2char)) if toks else odec.UNABLE
        assert formula == oformula, (formula, oformula)
        assert abs(conf - oconf) < 1e-5 * max(1.0, abs(conf)), (conf, oconf)
        print(f"api      image {i}: {len(toks)} tokens, confidence {conf:.6g}")
        out.append((formula, conf, lp))
    np.savez_compressed(os.path.join(GOLD, "swin_app_golden.npz"),
                        formulas=np.array([o[0] for o in out]),
                        confidences=np.array([o[1] for o in out], dtype=np.float64),
                        logprob_sums=np.array([o[2] for o in out], dtype=np.float64))
    print("wrote swin_app_golden.npz")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--phase", choices=["src", "app", "all"], default="all")
    a = ap.parse_args()
    if a.phase == "all":
        for ph in ("src", "app"):
            subprocess.check_call([sys.executable, "-m", "oracle.make_golden", "--phase", ph], cwd=ROOT)
    elif a.phase == "src":
        phase_src()
    else:
        phase_app()
