"""TemporalPDF facade (reference: src/temporal/pdf_compare.py:9-133): adaptive recursive pdf filter, 'l2' metric,
state kept on the device, update in one kernel (spef_pdf_filter)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from ..engine import Engine

_engine: Optional[Engine] = None


def _filter_engine() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine(32, 32, 8, 3, False, "fp32", 1)
    return _engine


class TemporalPDF:
    def __init__(self, n: float = 1.0, alpha: float = 1.0, distance_metric: str = 'l2'):
        if distance_metric.lower() != 'l2':
            raise NotImplementedError("only the 'l2' metric is implemented on the device (the only one Inference uses, "
                                      "src/temporal/inference.py:38-39)")
        self.n = n
        self.alpha = alpha
        self.distance_metric = 'l2'
        self._state: Optional[torch.Tensor] = None
        self._has: Optional[torch.Tensor] = None

    @property
    def previous_pdf(self) -> Optional[np.ndarray]:
        if self._state is None or int(self._has.item()) == 0:
            return None
        return self._state[0].cpu().numpy()

    def reset(self) -> None:
        """pdf_compare.py:24-30."""
        if self._has is not None:
            self._has.zero_()

    def update_pdf(self, current_pdf: np.ndarray) -> Tuple[np.ndarray, float]:
        """pdf_compare.py:94-133."""
        eng = _filter_engine()
        cur = torch.as_tensor(np.ascontiguousarray(current_pdf, np.float32)).reshape(1, -1).to(eng.device)
        n = cur.shape[1]
        if self._state is None or self._state.shape[1] != n:
            self._state = torch.zeros(1, n, dtype=torch.float32, device=eng.device)
            self._has = torch.zeros(1, dtype=torch.int32, device=eng.device)
        out, dist = eng.pdf_filter(cur, self._state, self._has, self.n, self.alpha)
        return out[0].cpu().numpy(), float(dist[0].item())
