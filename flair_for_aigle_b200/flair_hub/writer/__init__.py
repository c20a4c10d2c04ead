"""Mirror of the reference's ``flair_hub/writer`` for the part next to the hot path: the metric formulas applied to a
confusion matrix (``metrics_core.py``).  Prediction files, plots and the Lightning writer callback are out of scope."""
