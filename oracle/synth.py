"""Synthetic checkpoint / stroke images / vocab (re-exported from the product's workload
generator ``handwritten_math_ocr_api_b200.synthetic``: data definition, not arithmetic).
TEST INFRASTRUCTURE ONLY."""
from handwritten_math_ocr_api_b200.synthetic import (relative_position_index, state_dict_checksum,  # noqa: F401
                                                     synth_images, synth_state_dict, synth_stroke_image_u8,
                                                     synth_vocab)
