"""rmt_app_b200 — B200-native engine behind PyREMOT's `rmtExe` for the
pseudo-homogeneous packed-bed reactor models N1 (steady state) and N2
(dynamic, method of lines).  See DESIGN.md.

    from rmt_app_b200 import rmtExe, rmtCom, rmtExeBatch
"""
from .rmt import rmtExe, rmtCom, rmtExeBatch, rmtExeBatchN2      # noqa: F401
from .engine import solverSetting, Workspace       # noqa: F401

from .ensemble import rmtExeBatchSharded, rmtExeBatchN2Sharded            # noqa: F401

from .textkin import parse_reaction_rates           # noqa: F401
from .estimate import differential_evolution        # noqa: F401

__all__ = ["rmtExe", "rmtCom", "rmtExeBatch", "rmtExeBatchN2", "rmtExeBatchSharded", "rmtExeBatchN2Sharded", "solverSetting", "Workspace",
           "parse_reaction_rates", "differential_evolution"]
