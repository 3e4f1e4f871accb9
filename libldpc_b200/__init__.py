"""libldpc_b200 — B200-native LDPC decoder / Monte-Carlo simulator behind the heat1q/libldpc interfaces.

  libldpc_b200/libldpc.so   C ABI (include/ldpc_b200.h): the reference's six symbols + the handle API
  libldpc_b200/ldpcsim      CLI with the reference's flags
  libldpc_b200.ldpc.LDPC    drop-in for pyLDPC.ldpc.LDPC (same methods, same ctypes structs)
  libldpc_b200.api.Context  thin ctypes view of the handle API (batch decode, channel, sim rounds)
"""
from .api import Context, lib_path, load_library  # noqa: F401
from .ldpc import LDPC  # noqa: F401
