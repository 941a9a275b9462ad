"""Pin the ResNet-18 + TransformerEncoder oracle (oracle/res18_model.py) against the unmodified reference
(/root/reference/src/model_res18trans.py) and write tests/golden/res18_golden.npz.  Run HERE only.

    python oracle/make_golden_res18.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
SCRATCH = os.path.join(ROOT, "baseline", "_ref", "cwd")


def main():
    os.makedirs(SCRATCH, exist_ok=True)
    os.chdir(SCRATCH)                                  # config.py creates <cwd>/../checkpoints at import
    import torchvision
    orig = torchvision.models.resnet18
    torchvision.models.resnet18 = lambda weights=None, **k: orig(weights=None, **k)   # offline: no download
    from oracle import res18_model as R
    from oracle.arch import ModelConfig
    from handwritten_math_ocr_api_b200.layout import state_dict_layout_res18
    from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_pos_table, synth_state_dict_res18

    cfg = ModelConfig()
    sd = synth_state_dict_res18(cfg, seed=0)
    imgs = synth_images(4, seed=1234)
    pos = synth_pos_table(cfg.d_model, seed=0)

    sys.path.insert(0, os.path.join(REF, "src"))
    import model_res18trans as M                         # the reference, unmodified
    torch.manual_seed(0)
    model = M.FormulaRecognitionModel(cfg.vocab_size).eval()
    ref_sd = model.state_dict()
    manifest = [(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in ref_sd.items()]
    mine = [(k, list(s), d) for k, s, d in state_dict_layout_res18(cfg)]
    assert manifest == mine, "layout differs from the reference state_dict"
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected

    # the reference draws its positional table inside forward(): make nn.Embedding(10, d) return ours
    import torch.nn as nn
    real_embedding = nn.Embedding

    class FixedEmbedding(real_embedding):
        def __init__(self, n, d, *a, **k):
            super().__init__(n, d, *a, **k)
            if n == pos.shape[0] and d == pos.shape[1]:
                with torch.no_grad():
                    self.weight.copy_(pos)
    M.nn.Embedding = FixedEmbedding
    with torch.no_grad():
        trunk = model.encoder.features(imgs)
        d0 = (trunk - R.trunk(imgs, sd)).abs().max().item()
        print("trunk    max|ref-oracle| =", d0, " max|ref| =", trunk.abs().max().item())
        assert d0 < 1e-3 * max(1.0, trunk.abs().max().item())
        feats = model.encoder(imgs)                                   # [4,10,256]
        ofeats = R.encoder_forward(imgs, sd, cfg, pos)
        d1 = (feats - ofeats).abs().max().item()
        print("encoder  max|ref-oracle| =", d1, " max|ref| =", feats.abs().max().item(), feats.std().item())
        assert d1 < 5e-4, d1
        # D7: features depend on the batch composition
        alone = model.encoder(imgs[1:2])
        print("encoder  image 1 alone vs in the batch of 4: max diff", (alone - feats[1:2]).abs().max().item())
        g = torch.Generator().manual_seed(7)
        tgt = torch.randint(4, cfg.vocab_size, (4, 6), generator=g)
        tgt[:, 0] = cfg.sos
        logits = model.decoder(feats, tgt)
        ologits = R.decoder_forward(feats, tgt, sd, cfg)
        d2 = (logits - ologits).abs().max().item()
        print("decoder  max|ref-oracle| =", d2, " logits std =", logits.std().item())
        assert d2 < 5e-4, d2
        ys, step_logits = R.greedy_cached(ofeats, sd, cfg, max_len=40, return_logits=True)
        # the reference's own batched greedy loop restated over ITS modules (src/inference.py:15-25)
        ref_ys = torch.full((4, 1), cfg.sos, dtype=torch.long)
        fin = torch.zeros(4, dtype=torch.bool)
        for _ in range(40):
            nxt = model.decoder(feats, ref_ys)[:, -1, :].argmax(-1, keepdim=True)
            ref_ys = torch.cat([ref_ys, nxt], 1)
            fin |= nxt.squeeze(1) == cfg.eos
            if fin.all():
                break
        assert torch.equal(ref_ys, ys), "oracle greedy differs from the reference modules' greedy"
        top2 = step_logits.topk(2, -1).values
        margin = top2[..., 0] - top2[..., 1]
    np.savez_compressed(os.path.join(GOLD, "res18_golden.npz"), images_seed=np.int64(1234), weights_seed=np.int64(0),
                        pos_table=pos.numpy(), trunk=trunk.numpy(), features=feats.numpy(), tgt=tgt.numpy(),
                        logits=logits.numpy(), greedy_ys=ys.numpy(), greedy_margin=margin.numpy(),
                        versions=np.array(json.dumps({"torch": torch.__version__,
                                                      "torchvision": torchvision.__version__})))
    with open(os.path.join(GOLD, "res18_state_dict_manifest.json"), "w") as f:
        json.dump({"n_params": sum(p.numel() for p in model.parameters()), "entries": manifest}, f)
    print("wrote res18_golden.npz; greedy steps", ys.shape[1] - 1)


if __name__ == "__main__":
    main()
