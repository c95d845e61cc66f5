"""Batch-weighted running means (reference: src/tools/utils.py:16-104).  Host bookkeeping only."""
from __future__ import annotations

from typing import Dict, Tuple


class AverageMeter:
    """utils.py:67-104."""

    def __init__(self):
        self.reset()

    def reset(self) -> None:
        self.val = 0.0
        self.avg = 0.0
        self.sum = 0.0
        self.count = 0

    def update(self, val: float, n: int = 1) -> None:
        self.val = float(val)
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def get_avg(self) -> float:
        return self.avg


class RunningAverage:
    """utils.py:16-64."""

    def __init__(self, keys: Tuple[str, ...] = ('loss', 'accuracy')):
        self.running = {x: AverageMeter() for x in keys}

    def update(self, values: Dict[str, float], batch_size: int = 1) -> None:
        for key, value in values.items():
            self.running[key].update(value, batch_size)

    def get_multiple(self, keys: Tuple[str, ...] = ('loss', 'accuracy')) -> Dict[str, float]:
        return {x: self.running[x].get_avg() for x in keys}

    def get(self, key: str) -> float:
        assert key in self.running.keys(), f"Error: '{key}' not found in {self.running.keys()}"
        return self.running[key].get_avg()
