"""The reference's torch-dispatcher boundary (rlap/csrc/py_api_binder.cc:80-88), registered from Python:

    torch.ops.extension_cpp.approximate_cholesky(edge_info, num_nodes, num_remove, o_v, o_n) -> Tensor
    torch.ops.extension_cpp.identity(a) -> Tensor
    torch.ops.extension_cpp.approximate_cholesky_batched(edge_index, edge_weight, graph_ptr, num_remove, o_v, o_n,
                                                         num_views, seed) -> (Tensor edge_info, Tensor view_ptr)

The third op is new: many independent views, or the views of every graph of a batch, in one call;
edge_info is [sum E', 3] float64 with the views back to back, view_ptr int64 [num_views + 1] on the host.

Same namespace and schema strings; the kernels behind them are this package's CUDA path for CUDA tensors and, for
CPU tensors (what rlap/ops.py:47 hands over), the same path through the host-buffer C-ABI entry point
rlap_approximate_cholesky_host (copies inside). If the reference's own extension is loaded in the same process the
schemas already exist and only the CUDA kernels are added."""
import torch

from . import ops

_lib = None
_impl = None


def _edge_info_cuda(edge_info, num_nodes, num_remove, o_v, o_n):
    ei = edge_info[:, :2].t().long().contiguous()
    w = edge_info[:, 2].float().contiguous()
    return ops.approximate_cholesky(ei, w, int(num_nodes), int(num_remove), o_v, o_n)


def _edge_info_cpu(edge_info, num_nodes, num_remove, o_v, o_n):
    out = ops.approximate_cholesky_host(edge_info.detach().double().contiguous().numpy(), int(num_nodes),
                                        int(num_remove), o_v, o_n, seed=ops._next_seed())
    return torch.from_numpy(out)


def _batched(edge_index, edge_weight, graph_ptr, num_remove, o_v, o_n, num_views, seed):
    n = int(graph_ptr[-1])
    out, vp = ops.approximate_cholesky_batched(edge_index, edge_weight, n, num_remove.detach().cpu().numpy(), o_v, o_n,
                                               num_views=int(num_views), graph_ptr=graph_ptr, seed=int(seed))
    return out, vp


def register():
    """idempotent; returns True if this call (or an earlier one) registered the kernels"""
    global _lib, _impl
    if _impl is not None:
        return True
    try:
        _lib = torch.library.Library("extension_cpp", "DEF")
        _lib.define("approximate_cholesky(Tensor edge_info, int num_nodes, int num_remove, str o_v,  str o_n) -> Tensor")
        _lib.define("identity(Tensor a) -> Tensor")
        own_schema = True
    except RuntimeError:
        own_schema = False      # the reference's extension already defined the namespace
        _lib = torch.library.Library("extension_cpp", "FRAGMENT")
    _lib.define("approximate_cholesky_batched(Tensor edge_index, Tensor? edge_weight, Tensor graph_ptr, Tensor num_remove, "
                "str o_v, str o_n, int num_views, int seed) -> (Tensor, Tensor)")
    _impl = torch.library.Library("extension_cpp", "IMPL")
    _impl.impl("approximate_cholesky_batched", _batched, "CUDA")
    _impl.impl("approximate_cholesky_batched", _batched, "CPU")
    _impl.impl("approximate_cholesky", _edge_info_cuda, "CUDA")
    _impl.impl("identity", lambda a: a.clone(), "CUDA")
    if own_schema:
        _impl.impl("approximate_cholesky", _edge_info_cpu, "CPU")
        _impl.impl("identity", lambda a: a.clone(), "CPU")
    return True
