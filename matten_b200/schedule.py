"""Learning-rate schedule and stopping rule of the reference's training configs (host-side control logic):
``torch.optim.lr_scheduler.ReduceLROnPlateau(mode="min", factor=0.5, patience=50)`` and
``pytorch_lightning.callbacks.EarlyStopping(monitor="val/score", mode="min", patience=150, min_delta=0)``
(scripts/configs/materials_tensor.yaml:70-95, pretrained/20230627/config_final.yaml:17-23) for ``train.FlatAdam``,
which is not a ``torch.optim.Optimizer``."""
from __future__ import annotations

import math


class ReduceLROnPlateau:
    """Same update rule as torch's scheduler (threshold_mode "rel", cooldown 0): after ``patience`` consecutive epochs
    without an improvement by more than ``threshold`` (relative), multiply the learning rate by ``factor``."""

    def __init__(self, optimizer, mode: str = "min", factor: float = 0.5, patience: int = 50, threshold: float = 1e-4,
                 min_lr: float = 0.0, eps: float = 1e-8):
        if factor >= 1.0:
            raise ValueError("Factor should be < 1.0.")
        if mode not in ("min", "max"):
            raise ValueError(f"mode {mode} is unknown!")
        self.optimizer, self.mode, self.factor, self.patience = optimizer, mode, factor, patience
        self.threshold, self.min_lr, self.eps = threshold, min_lr, eps
        self.best = math.inf if mode == "min" else -math.inf
        self.num_bad_epochs = 0

    def _is_better(self, a: float) -> bool:
        if self.mode == "min":
            return a < self.best * (1.0 - self.threshold)
        return a > self.best * (1.0 + self.threshold)

    def step(self, metric: float) -> float:
        metric = float(metric)
        if self._is_better(metric):
            self.best = metric
            self.num_bad_epochs = 0
        else:
            self.num_bad_epochs += 1
        if self.num_bad_epochs > self.patience:
            new_lr = max(self.optimizer.lr * self.factor, self.min_lr)
            if self.optimizer.lr - new_lr > self.eps:
                self.optimizer.lr = new_lr
            self.num_bad_epochs = 0
        return self.optimizer.lr


class EarlyStopping:
    """``should_stop`` turns True after ``patience`` checks without an improvement of at least ``min_delta``."""

    def __init__(self, mode: str = "min", patience: int = 150, min_delta: float = 0.0):
        self.mode, self.patience, self.min_delta = mode, patience, abs(min_delta)
        self.best = math.inf if mode == "min" else -math.inf
        self.wait = 0
        self.should_stop = False

    def step(self, metric: float) -> bool:
        metric = float(metric)
        improved = (metric < self.best - self.min_delta) if self.mode == "min" else (metric > self.best + self.min_delta)
        if improved:
            self.best, self.wait = metric, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.should_stop = True
        return self.should_stop
