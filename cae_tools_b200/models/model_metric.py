"""Post-training metrics (mse / rmse / mae / mean per-case Pearson) over masked pixels.
Host-side numpy; mirrors what the reference reports at the end of ``train``
(reference: src/cae_tools/models/model_metric.py:19-71)."""

import numpy as np


class ModelMetric:

    def __init__(self):
        self.sq = 0.0
        self.ab = 0.0
        self.count = 0
        self.correlations = []

    def accumulate(self, actual, estimates, mask):
        if actual.shape != estimates.shape:
            raise ValueError("The shapes of 'actual' and 'estimates' must match.")
        keep = np.broadcast_to(np.asarray(mask).astype(bool), actual.shape).reshape(-1)
        a = np.asarray(actual, dtype=np.float64).reshape(-1)[keep]
        e = np.asarray(estimates, dtype=np.float64).reshape(-1)[keep]
        d = a - e
        self.sq += float(np.sum(d * d))
        self.ab += float(np.sum(np.abs(d)))
        self.count += a.size
        if a.size:
            ac, ec = a - a.mean(), e - e.mean()
            den = np.sqrt(np.sum(ac * ac) * np.sum(ec * ec))
            self.correlations.append(float(np.sum(ac * ec) / den) if den > 0 else float("nan"))

    def get_metrics(self):
        if self.count == 0:
            raise ValueError("No data accumulated to calculate metrics.")
        mse = self.sq / self.count
        return {
            "mse": mse,
            "rmse": float(np.sqrt(mse)),
            "mae": self.ab / self.count,
            "mean_pearson_correlation": float(np.mean(self.correlations)) if self.correlations else 0.0,
        }
