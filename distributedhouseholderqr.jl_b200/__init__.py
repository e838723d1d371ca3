"""dhqr_b200 — B200-native blocked Householder QR behind DistributedHouseholderQR.jl's qr! / \\.

Import as ``import dhqr_b200`` (repo-root shim) — the directory keeps the name the task fixes
(``distributedhouseholderqr.jl_b200``), which is not a valid Python identifier.
"""
from . import _lib
from .api import (ColumnBlockMatrix, DistributedHouseholderQRStruct, Handle, LocalColumnBlock, alphafactor,
                  apply_q_, apply_qt_, backsolve_, balanced_splits, colmajor_empty, default_handle, fill_uniform_, householder_, init_distributed, ldiv,
                  partialdot, plan_host_upload, qr_, qr_bang, shutdown_distributed, solve_householder_, splits, to_colmajor)

__all__ = ["ColumnBlockMatrix", "DistributedHouseholderQRStruct", "Handle", "LocalColumnBlock", "alphafactor",
           "apply_q_", "apply_qt_", "backsolve_", "balanced_splits", "colmajor_empty", "default_handle", "fill_uniform_", "householder_", "init_distributed",
           "ldiv", "partialdot", "plan_host_upload", "qr_", "qr_bang", "shutdown_distributed", "solve_householder_", "splits",
           "to_colmajor", "_lib"]
