"""B200-native batched TinyMPC (cached-Riccati ADMM) -- Python face of the C ABI in include/tmpc.h.

The directory name carries a hyphen, so import it through `__graft_entry__.load_package()` (which
registers it as `accelerated_tinympc_b200`).  PyTorch is plumbing only (device buffers, streams,
torch.distributed); every numerical result comes from the hand-written sm_100a kernels in csrc/.
"""
from . import capi, problems, sharding, workloads  # noqa: F401

__all__ = ["capi", "problems", "sharding", "workloads"]
