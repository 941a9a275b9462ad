"""GPU parity of the ResNet-18 + TransformerEncoder variant (BASELINE.json config 4,
/root/reference/src/model_res18trans.py) against golden vectors made by the unmodified reference
(oracle/make_golden_res18.py) and against the oracle on the same inputs.

Tolerances (fp16 operands, fp32 accumulation; trunk activations reach ~330 with the synthetic He-scaled
checkpoint, encoder features have std ~1.0): features max-abs <= 1.5e-2 (measured 1.0e-2), teacher-forced logits
<= 1e-2 (north_star's figure; measured 7.7e-3), greedy tokens identical or diverging at a reference margin < 2e-2.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FEAT_TOL = 1.5e-2
LOGIT_TOL = 1e-2


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "res18_golden.npz"))


@pytest.fixture(scope="module")
def rsd(cfg):
    from handwritten_math_ocr_api_b200.synthetic import synth_state_dict_res18
    return synth_state_dict_res18(cfg, seed=0)


@pytest.fixture(scope="module")
def rmodel(rsd, cfg):
    from handwritten_math_ocr_api_b200.model_res18trans import FormulaRecognitionModel
    m = FormulaRecognitionModel(cfg.vocab_size)
    m.load_state_dict(rsd)
    return m.eval()


def _imgs(gold, n=4):
    from handwritten_math_ocr_api_b200.synthetic import synth_images
    return synth_images(n, int(gold["images_seed"]))


def test_surface_and_param_count(rmodel, rsd):
    n = sum(v.numel() for k, v in rsd.items()
            if k != "decoder.tgt_mask" and not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert sum(p.numel() for p in rmodel.parameters()) == n
    with pytest.raises(RuntimeError):
        rmodel.to("cpu")


def test_encoder_against_reference_golden(rmodel, gold):
    pos = torch.from_numpy(gold["pos_table"])
    feats = rmodel.encoder(_imgs(gold).cuda(), pos)
    assert feats.shape == (4, 10, 256) and feats.dtype == torch.float32
    err = np.abs(feats.cpu().numpy() - gold["features"]).max()
    print("res18 encoder max-abs error vs reference:", err, "feature std", gold["features"].std())
    assert err < FEAT_TOL


def test_encoder_attends_across_the_batch(rmodel, gold):
    """batch_first=True fed [10, B, d] - an image's features depend on its batch-mates."""
    pos = torch.from_numpy(gold["pos_table"])
    imgs = _imgs(gold).cuda()
    full = rmodel.encoder(imgs, pos)
    alone = rmodel.encoder(imgs[1:2], pos)
    assert (alone - full[1:2]).abs().max().item() > 1e-2
    # and a fresh positional table is drawn per call when none is given (model_res18trans.py:57-59)
    a, b = rmodel.encoder(imgs), rmodel.encoder(imgs)
    assert (a - b).abs().max().item() > 1e-2


def test_decoder_against_reference_golden(rmodel, gold):
    logits = rmodel.decoder(torch.from_numpy(gold["features"]).cuda(), torch.from_numpy(gold["tgt"]).cuda())
    assert logits.shape == (4, 6, 5075)
    err = np.abs(logits.cpu().numpy() - gold["logits"]).max()
    print("res18 decoder max-abs logit error vs reference:", err)
    assert err < LOGIT_TOL


def test_greedy_generate_against_reference_golden(rmodel, gold, cfg):
    pos = torch.from_numpy(gold["pos_table"])
    tokens, steps, logp = rmodel.generate(_imgs(gold).cuda(), max_len=40, return_logprobs=True, pos_table=pos)
    ref = gold["greedy_ys"]
    got = tokens.cpu().numpy()
    margin = gold["greedy_margin"]
    same = 0
    for r in range(4):
        n = min(got.shape[1], ref.shape[1])
        d = np.nonzero(got[r, :n] != ref[r, :n])[0]
        if d.size == 0:
            same += 1
            continue
        c = int(d[0])
        print(f"row {r}: first divergence at column {c}, reference margin {float(margin[r, c - 1]):.4f}")
        assert float(margin[r, c - 1]) < 2 * LOGIT_TOL
    print(f"res18 identical greedy sequences: {same}/4, steps {steps} (reference {ref.shape[1] - 1})")


def test_beam_search_on_the_variant_scores_like_the_oracle(rmodel, rsd, gold, cfg):
    """Beam 3 through the 10-token memory: the reported score is the oracle's log-probability of the returned
    sequence (scoring + back-track), and it is never worse than the greedy sequence's."""
    from oracle import res18_model as R
    pos = torch.from_numpy(gold["pos_table"])
    imgs = _imgs(gold).cuda()
    feats = rmodel.encoder(imgs, pos)
    tokens, steps, _, score = rmodel.generate(encoder_out=feats, max_len=20, beam_size=3)
    g_tok, _, g_lp = rmodel.generate(encoder_out=feats, max_len=20, return_logprobs=True)
    assert tokens.shape == (4, steps + 1) and torch.isfinite(score).all()

    def oracle_score(tok):
        with torch.no_grad():
            lsm = torch.log_softmax(R.decoder_forward(feats.cpu(), tok[:, :-1].cpu(), rsd, cfg).float(), -1)
        out = []
        for r in range(tok.shape[0]):
            s = 0.0
            for t in range(1, tok.shape[1]):
                tk = int(tok[r, t])
                s += float(lsm[r, t - 1, tk])
                if tk == cfg.eos:
                    break
            out.append(s)
        return torch.tensor(out)

    assert (oracle_score(tokens) - score.cpu()).abs().max().item() < 3e-2
    assert (score.cpu() >= oracle_score(g_tok) - 3e-2).all()


@pytest.mark.parametrize("batch", [1, 3, 4, 9])
def test_implicit_gemm_convolutions_match_the_im2col_path(rmodel, gold, batch):
    """conv_impl 0 (4-D TMA patches feeding the tcgen05 GEMM, zero padding = out-of-bounds fill, stride =
    elementStrides) against conv_impl 1 (explicit im2col matrix + the same GEMM): same fp16 operands, same fp32
    accumulation order - bit-identical features.  Batches that do not fill the last 4-image tile of layer4 included."""
    from handwritten_math_ocr_api_b200.synthetic import synth_images
    imgs = synth_images(batch, 77).cuda()
    pos = torch.from_numpy(gold["pos_table"])
    try:
        rmodel.set_option("conv_impl", 1)
        ref = rmodel.encoder(imgs, pos).clone()
    finally:
        rmodel.set_option("conv_impl", 0)
    out = rmodel.encoder(imgs, pos)
    assert torch.equal(out, ref)        # same fp16 operands, same K order (tap-major, then channel): the same bits
