"""bigkrls_b200 - B200-native estimation hot path of bigKRLS behind the reference's API.

Product code path: bigkrls_b200.api (host glue) -> bigkrls_b200._lib (ctypes) ->
libbigkrls_b200.so (hand-written sm_100a CUDA).  Nothing here imports `oracle/`.
"""
from .api import BigKRLS, bigKRLS, crossvalidate_bigKRLS, predict, summary  # noqa: F401
from ._lib import BKError, Context, default_context  # noqa: F401

__all__ = ["bigKRLS", "predict", "summary", "crossvalidate_bigKRLS", "BigKRLS", "BKError", "Context",
           "default_context"]
