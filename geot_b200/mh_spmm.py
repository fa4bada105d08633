"""``geot.mh_spmm`` / ``mh_spmm_transposed`` (reference: ``geot/mh_spmm.py:4-12``)."""
import torch


def mh_spmm(src_index: torch.Tensor, dst_index: torch.Tensor, weight: torch.Tensor, src: torch.Tensor,
            reduce: str = "sum") -> torch.Tensor:
    """``out[dst[e], h, :] (reduce)= weight[e, h] * src[src[e], h, :]``; src ``[N, H, F]``.

    ``weight`` is ``[E, H]`` or ``[H, E]``; the layout is read from the shape, ``[E, H]`` first
    (``csrc/cuda/wrapper/mh_spmm_base.h:38-49``).  fp32 / fp64 / bf16 / fp16 (fp32 accumulation).
    """
    return torch.ops.geot.mh_spmm(src_index, dst_index, weight, src, reduce)


def mh_spmm_transposed(src_index: torch.Tensor, dst_index: torch.Tensor, weight: torch.Tensor,
                       src: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
    """Takes ``weight [E, H]`` and runs the ``[H, E]`` layout, like the reference
    (``geot/mh_spmm.py:8-12``)."""
    weight = weight.transpose(0, 1).contiguous()
    return torch.ops.geot.mh_spmm(src_index, dst_index, weight, src, reduce)
