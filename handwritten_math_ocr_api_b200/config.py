"""Hyper-parameters of the hot path, same attribute names as the reference's ``config`` singletons
(``/root/reference/src/config.py:16-50`` and ``/root/reference/app/src/config.py:22-57``).

The engine never keeps its own copy of a hyper-parameter: ``FormulaRecognitionModel`` reads
whatever object it is given (this default, or the reference's own ``config``).
"""


class Config:
    img_w = 320
    img_h = 96
    d_model = 256
    nhead = 8
    dim_feedforward = 512
    dropout = 0.2                      # identity in eval; the engine is inference-only
    swin_num_decoder_layers = 8        # src/config.py:32
    num_decoder_layers = 8             # app/src/config.py:27
    res18trans_num_encoder_layers = 8  # src/config.py:28
    res18trans_num_decoder_layers = 8  # src/config.py:29
    max_seq_len = 150
    sos_token = '<sos>'
    eos_token = '<eos>'
    pad_token = '<pad>'
    unk_token = '<unk>'
    special_tokens = [pad_token, sos_token, eos_token, unk_token]
    beam_size = 5                      # src/config.py:50 (unused by the reference's greedy loops)


config = Config()
