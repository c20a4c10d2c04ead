"""Parity bars shared by the GPU tests (one place to state them; DESIGN.md section 2 quotes these).

BASELINE.json north_star: "logits must match the reference's own PyTorch path within bf16 tolerance (max-abs and
relative error stated), and the class map must agree on at least 99.9% of pixels".
"""
# fraction of ALL pixels of a zone / tile whose class equals the fp32 reference's (raw, no confidence mask)
CLASS_AGREEMENT = 0.98
# logit error bars, as fractions of the logit standard deviation of the tile
LOGIT_MEAN_ABS = 0.015
LOGIT_MAX_ABS = 0.15
