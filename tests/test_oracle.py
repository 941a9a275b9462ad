"""CPU: the oracle restatement against the golden vectors produced by the real reference
(``oracle/make_golden.py``) and the reference's one published known answer (parameter count)."""
import json
import os

import numpy as np
import torch

from oracle import decode as odec
from oracle import ref_model as oref
from oracle.arch import ModelConfig, stage_dims, state_dict_layout
from oracle.synth import (relative_position_index, state_dict_checksum, synth_images,
                          synth_state_dict, synth_vocab)

from conftest import GOLDEN


def test_layout_matches_reference_manifest(cfg):
    man = json.load(open(os.path.join(GOLDEN, "state_dict_manifest.json")))
    mine = [[k, list(s), d] for k, s, d in state_dict_layout(cfg)]
    assert mine == man["entries"]
    assert len(mine) == 517


def test_param_count_known_answer(cfg, sd):
    # README.md:89 / images/mlflow-swin.png: 37.45 M parameters at V=5075
    n = sum(v.numel() for k, v in sd.items()
            if not k.startswith("encoder.features.") and "relative_position_index" not in k
            and k != "decoder.tgt_mask")
    assert n == 37_450_293


def test_checkpoint_is_reproducible(sd, golden_src):
    assert state_dict_checksum(sd) == str(golden_src["checksum"])


def test_relative_position_index_formula():
    # torchvision builds it with meshgrid arithmetic; ours is the closed form of SURVEY.md appendix A
    idx = relative_position_index().reshape(49, 49)
    assert idx[0, 0] == 6 * 13 + 6 and idx[0, 48] == 0 and idx[48, 0] == 168


def test_stage_dims():
    assert stage_dims() == [(24, 80, 96, 3), (12, 40, 192, 6), (6, 20, 384, 12), (3, 10, 768, 24)]


def test_window_geometry_regions():
    # stage 3 (6x20 -> 7x21, shift (0,3)) has only ids {6,7,8} in 3 bands
    Hp, Wp, sh, sw, src, valid, region = oref.window_geometry(6, 20, 3)
    assert (Hp, Wp, sh, sw) == (7, 21, 0, 3)
    rows = region.reshape(1, 3, 7, 7).permute(0, 2, 1, 3).reshape(7, 21)
    assert rows[0].tolist() == [6] * 14 + [7] * 4 + [8] * 3
    assert all(rows[i].tolist() == rows[0].tolist() for i in range(7))
    # stage 4 pads 3x10 -> 7x14
    Hp, Wp, sh, sw, src, valid, region = oref.window_geometry(3, 10, 3)
    assert (Hp, Wp, sh, sw) == (7, 14, 0, 3) and int(valid.sum()) == 30
    # stage 1 uses all nine ids
    Hp, Wp, sh, sw, src, valid, region = oref.window_geometry(24, 80, 3)
    assert (Hp, Wp, sh, sw) == (28, 84, 3, 3) and sorted(region.unique().tolist()) == list(range(9))
    assert int(valid.sum()) == 24 * 80 and sorted(src[valid].tolist()) == list(range(24 * 80))


def test_images_are_deterministic(golden_src):
    a, b = synth_images(3, int(golden_src["images_seed"])), synth_images(3, int(golden_src["images_seed"]))
    assert torch.equal(a, b) and a.shape == (3, 1, 96, 320)
    assert a.min() == -1.0 and a.max() == 1.0


def test_encoder_against_reference_golden(sd, golden_src):
    imgs = synth_images(4, int(golden_src["images_seed"]))
    with torch.no_grad():
        feats = oref.encoder_forward(imgs, sd)
    assert feats.shape == (4, 30, 256)
    assert np.abs(feats.numpy() - golden_src["features"]).max() < 2e-4


def test_decoder_against_reference_golden(sd, cfg, golden_src):
    feats = torch.from_numpy(golden_src["features"][:2])
    tgt = torch.from_numpy(golden_src["tgt"])
    with torch.no_grad():
        logits = oref.decoder_forward(feats, tgt, sd, cfg)
    assert logits.shape == (2, 6, cfg.vocab_size)
    assert np.abs(logits.numpy() - golden_src["logits"]).max() < 5e-4


def test_cached_greedy_reproduces_reference_inference(sd, cfg, golden_src):
    """``/root/reference/src/inference.py::predict`` (run unmodified when the fixture was made)."""
    feats = torch.from_numpy(golden_src["features"])
    with torch.no_grad():
        ys = odec.greedy_cached(feats, sd, cfg)
    assert np.array_equal(ys.numpy(), golden_src["greedy_ys"])
    _, idx2char = synth_vocab(cfg.vocab_size)
    strings = [odec.ids_to_string(r.tolist(), idx2char, cfg) for r in ys]
    assert strings == [str(s) for s in golden_src["greedy_strings"]]
    # early stop rule: all four never finish together before 150 steps in this fixture
    assert ys.shape[1] == 151


def test_early_exit_when_all_finished(sd, cfg, golden_src):
    feats = torch.from_numpy(golden_src["features"][1:3])       # both emit EOS at step 19
    with torch.no_grad():
        ys = odec.greedy_cached(feats, sd, cfg)
    assert ys.shape[1] == 20 and (ys[:, -1] == cfg.eos).all()
    assert np.array_equal(ys.numpy(), golden_src["greedy_ys"][1:3, :20])


def test_predict_py_semantics(sd, cfg, golden_src):
    # predict.py:67 returns output_seq[1:-1]: for image 1 (EOS at step 19) that is 18 tokens
    toks = golden_src["greedy_ys"][1, 1:19].tolist()
    feats = torch.from_numpy(golden_src["features"][1:2])
    with torch.no_grad():
        ys = odec.greedy_cached(feats, sd, cfg)
    assert ys[0, 1:-1].tolist() == toks
    assert golden_src["single_tokens"].shape[0] == 149        # image 0: length-terminated, last token dropped


def test_api_confidence_against_reference_golden(sd, cfg, golden_app, golden_src):
    """``/root/reference/app/src/im2latex.py::predict`` (run unmodified when the fixture was made)."""
    _, idx2char = synth_vocab(cfg.vocab_size)
    img = synth_images(2, int(golden_src["images_seed"]))[1:2]
    toks, lp, conf = odec.greedy_api(img, sd, cfg)
    assert odec.clean_latex_output(odec.tokens_to_latex(toks, idx2char)) == str(golden_app["formulas"][1])
    assert abs(conf - float(golden_app["confidences"][1])) < 1e-5
    assert abs(lp - float(golden_app["logprob_sums"][1])) < 1e-3


def test_beam_invariants(sd, cfg, golden_src):
    feats = torch.from_numpy(golden_src["features"][:3])
    with torch.no_grad():
        g = odec.greedy_cached(feats, sd, cfg, max_len=24)
        t1, s1, _, _ = odec.beam_search(feats, sd, cfg, beam=1, max_len=24)
        t3, s3, all3, sc3 = odec.beam_search(feats, sd, cfg, beam=3, max_len=24)
        t3b, s3b, _, _ = odec.beam_search(feats[1:2], sd, cfg, beam=3, max_len=24)
    _, idx2char = synth_vocab(cfg.vocab_size)
    to_s = lambda r: odec.ids_to_string(r.tolist(), idx2char, cfg)
    assert [to_s(r) for r in t1] == [to_s(r) for r in g]         # beam=1 == greedy up to EOS
    assert (s3 >= s1 - 1e-4).all()                               # wider beam never scores lower
    assert to_s(t3b[0]) == to_s(t3[1]) and abs(float(s3b[0] - s3[1])) < 1e-4   # batch invariance
    assert (sc3[:, :-1] >= sc3[:, 1:]).all()                     # hypotheses come out sorted


def test_clean_latex_output():
    assert odec.clean_latex_output(r"\begin {matrix} a \end {matrix}") == r"\begin{matrix} a \end{matrix}"
    assert odec.clean_latex_output(r"\mathrm { abc }") == r"\mathrm {abc}"


# ---- ResNet-18 + TransformerEncoder variant (BASELINE.json config 4) --------------------------------------
def test_res18_oracle_against_reference_golden():
    """oracle/res18_model.py reproduces the outputs the unmodified /root/reference/src/model_res18trans.py
    produced for the same synthetic checkpoint, images and positional table (oracle/make_golden_res18.py)."""
    import json
    import os

    import numpy as np
    import torch

    from handwritten_math_ocr_api_b200.layout import ModelConfig, state_dict_layout_res18
    from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_pos_table, synth_state_dict_res18
    from oracle import res18_model as R

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "res18_golden.npz"))
    cfg = ModelConfig()
    sd = synth_state_dict_res18(cfg, seed=int(gold["weights_seed"]))
    manifest = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "res18_state_dict_manifest.json")))
    assert [(k, list(s), d) for k, s, d in state_dict_layout_res18(cfg)] == [tuple(e) if False else (e[0], e[1], e[2])
                                                                            for e in manifest["entries"]]
    assert len(manifest["entries"]) == 367
    imgs = synth_images(4, int(gold["images_seed"]))
    pos = synth_pos_table(cfg.d_model, seed=0)
    assert np.array_equal(pos.numpy(), gold["pos_table"])
    with torch.no_grad():
        assert np.abs(R.trunk(imgs, sd).numpy() - gold["trunk"]).max() < 1e-2          # activations up to ~330
        feats = R.encoder_forward(imgs, sd, cfg, pos)
        assert np.abs(feats.numpy() - gold["features"]).max() < 1e-4
        logits = R.decoder_forward(torch.from_numpy(gold["features"]), torch.from_numpy(gold["tgt"]), sd, cfg)
        assert np.abs(logits.numpy() - gold["logits"]).max() < 1e-3
        ys = R.greedy_cached(torch.from_numpy(gold["features"]), sd, cfg, max_len=40)
        assert np.array_equal(ys.numpy(), gold["greedy_ys"])
        # the encoder attends across the batch - an image's features depend on its batch-mates
        alone = R.encoder_forward(imgs[1:2], sd, cfg, pos)
        assert (alone - feats[1:2]).abs().max().item() > 1e-2


def test_preprocess_oracle_matches_pil_and_torchvision():
    """oracle/preprocess.py (restated Pillow convert + resample, torchvision ToTensor + Normalize) against the real
    libraries run on the same pixels: the transform of app/src/preprocess.py:6-16, bit for bit, for up- and
    down-scaling in either axis, both input modes and degenerate sizes."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    from oracle.preprocess import reference_preprocess
    t = transforms.Compose([transforms.Grayscale(num_output_channels=1), transforms.Resize((96, 320)),
                            transforms.ToTensor(), transforms.Normalize(mean=[0.5], std=[0.5])])
    rng = np.random.default_rng(5)
    sizes = [(96, 320), (48, 160), (200, 800), (97, 321), (1, 1), (3, 1000), (500, 7), (640, 480), (95, 319), (31, 333)]
    for i, (h, w) in enumerate(sizes):
        for mode in ("L", "RGB"):
            arr = rng.integers(0, 256, (h, w) if mode == "L" else (h, w, 3), dtype=np.uint8)
            want = t(Image.fromarray(arr, mode=mode)).unsqueeze(0).numpy()
            got = reference_preprocess(arr)
            assert got.shape == want.shape == (1, 1, 96, 320)
            assert np.array_equal(got, want), (h, w, mode, np.abs(got - want).max())


def test_dataloader_preprocess_oracle_matches_cv2():
    """oracle.preprocess.cv2_resize_linear / reference_preprocess_dataloader against the loader's own calls
    (src/data_loader.py:31-35: cv2.resize(img, (320, 96)) -> PIL "L" -> the transform), bit for bit."""
    import cv2
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    from oracle.preprocess import cv2_resize_linear, reference_preprocess_dataloader
    t = transforms.Compose([transforms.Grayscale(num_output_channels=1), transforms.Resize((96, 320)),
                            transforms.ToTensor(), transforms.Normalize(mean=[0.5], std=[0.5])])
    rng = np.random.default_rng(3)
    for h, w in [(96, 320), (48, 160), (192, 640), (200, 800), (97, 321), (50, 170), (500, 1300), (33, 77), (95, 319),
                 (1, 1), (2, 5), (96, 1000), (300, 320), (7, 4000), (193, 641), (24, 80)]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        r = cv2.resize(a, (320, 96))
        assert np.array_equal(cv2_resize_linear(a, 96, 320), r), (h, w)
        want = t(Image.fromarray(r, mode='L')).numpy()
        assert np.array_equal(reference_preprocess_dataloader(a), want), (h, w)
