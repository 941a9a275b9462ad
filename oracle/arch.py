"""Architecture constants / state-dict layout (re-exported from the product's pure-Python
``layout`` module so the oracle, the golden-vector generator and the engine agree on one
definition; the layout itself is checked against the real reference's ``state_dict()`` by
``oracle/make_golden.py`` and ``tests/test_oracle.py``).  TEST INFRASTRUCTURE ONLY."""
from handwritten_math_ocr_api_b200.layout import *  # noqa: F401,F403
from handwritten_math_ocr_api_b200.layout import ModelConfig, stage_dims, state_dict_layout  # noqa: F401
