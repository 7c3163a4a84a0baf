"""distegnn_b200 — B200-native (sm_100a) implementation of the DistEGNN hot path.

Public surface mirrors the reference: ``from distegnn_b200 import FastEGNN`` is a drop-in for
``from models.FastEGNN import FastEGNN``.
"""
from .fast_egnn import E_GCL_vel, FastEGNN  # noqa: F401

__all__ = ["FastEGNN", "E_GCL_vel"]
