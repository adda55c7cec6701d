"""NVTX ranges for the stages of the hot path (the reference has no tracing hooks at all)."""
import contextlib

import torch

from . import config


@contextlib.contextmanager
def nvtx(name):
    on = bool(config.NVTX) and torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()
