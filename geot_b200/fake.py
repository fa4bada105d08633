"""Fake (meta) implementations of the C++-registered ``geot::*`` operators so that ``torch.export`` /
``torch.compile`` can trace through them.  The reference registers fakes only for its Python custom ops
(``geot/gather_scatter.py:12-18``, ``geot/gather_weight_scatter.py:21-28``, ``geot/csr_gws.py:30-37``);
``geot::mh_spmm`` and ``geot::index_scatter`` have none there.  The number of output rows is
data dependent (``index[-1] + 1``), hence a fresh dynamic size.
"""
import torch


def _rows():
    return torch.library.get_ctx().new_dynamic_size()


@torch.library.register_fake("geot::index_scatter")
def _(dim, index, src, reduce, sorted):
    shape = list(src.shape)
    shape[dim] = _rows()
    return src.new_empty(shape)


@torch.library.register_fake("geot::gather_scatter_impl")
def _(src_index, dst_index, src):
    return src.new_empty([_rows(), src.shape[1]])


@torch.library.register_fake("geot::gather_scatter_reduce")
def _(src_index, dst_index, src, reduce):
    return src.new_empty([_rows(), src.shape[1]])


@torch.library.register_fake("geot::gather_weight_scatter_impl")
def _(src_index, dst_index, weight, src):
    return src.new_empty([_rows(), src.shape[1]])


@torch.library.register_fake("geot::gather_weight_scatter_reduce")
def _(src_index, dst_index, weight, src, reduce):
    return src.new_empty([_rows(), src.shape[1]])


@torch.library.register_fake("geot::mh_spmm")
def _(src_index, dst_index, weight, src, reduce):
    return src.new_empty([_rows(), src.shape[1], src.shape[2]])


@torch.library.register_fake("geot::sddmm_coo_impl")
def _(src_index, dst_index, mat_1, mat_2):
    return mat_1.new_empty([src_index.shape[0]])


@torch.library.register_fake("geot::csr_gws_impl")
def _(indptr, indices, weight, src):
    return src.new_empty([indptr.shape[0], src.shape[1]])


@torch.library.register_fake("geot::coo_to_csr_impl")
def _(coo_row):
    return coo_row.new_empty([_rows()], dtype=torch.int32)
