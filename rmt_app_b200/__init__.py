"""rmt_app_b200 — B200-native engine behind PyREMOT's `rmtExe` for the
pseudo-homogeneous packed-bed reactor models N1 (steady state) and N2
(dynamic, method of lines).  See DESIGN.md.

    from rmt_app_b200 import rmtExe, rmtCom, rmtExeBatch
"""
from .rmt import rmtExe, rmtCom, rmtExeBatch      # noqa: F401
from .engine import solverSetting, Workspace       # noqa: F401

from .ensemble import rmtExeBatchSharded            # noqa: F401

__all__ = ["rmtExe", "rmtCom", "rmtExeBatch", "rmtExeBatchSharded", "solverSetting", "Workspace"]
