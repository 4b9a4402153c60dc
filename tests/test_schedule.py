"""Host-side training control: the plateau scheduler follows torch's on the same metric sequence; early stopping."""
import random

import torch

from matten_b200.schedule import EarlyStopping, ReduceLROnPlateau


class _Opt:
    def __init__(self, lr):
        self.lr = lr


def test_reduce_lr_on_plateau_matches_torch():
    rng = random.Random(0)
    seq = [1.0 / (1 + 0.05 * i) + 0.02 * rng.random() for i in range(40)] + [0.4 + 0.01 * rng.random() for _ in range(60)]
    p = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.Adam([p], lr=0.01)
    tsch = torch.optim.lr_scheduler.ReduceLROnPlateau(topt, mode="min", factor=0.5, patience=5)
    mine = _Opt(0.01)
    msch = ReduceLROnPlateau(mine, mode="min", factor=0.5, patience=5)
    for v in seq:
        tsch.step(v)
        msch.step(v)
        assert abs(topt.param_groups[0]["lr"] - mine.lr) < 1e-15
    assert mine.lr < 0.01  # the plateau was detected


def test_early_stopping():
    es = EarlyStopping(mode="min", patience=3)
    flags = [es.step(v) for v in [1.0, 0.9, 0.95, 0.91, 0.9, 0.85]]
    assert flags == [False, False, False, False, True, True]
    es = EarlyStopping(mode="min", patience=3)
    assert not any(es.step(v) for v in [1.0, 0.9, 0.95, 0.8, 0.85, 0.7])
