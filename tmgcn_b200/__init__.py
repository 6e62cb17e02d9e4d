"""tmgcn_b200 -- B200 (sm_100a) implementation of the TM-GCN propagation hot path.

Host-side mirror of the reference's call surface (`func_MProduct`,
`create_matrix_M`, `EmbeddingGCN`, `EmbeddingGCN2`, `EmbeddingKWGCN`) over the
C ABI of libtmgcn_b200.so (include/tmgcn.h).  Importing the package does not
need a GPU; calling anything on the hot path does, and fails loudly without one.
"""
from . import _lib  # noqa: F401
from .modules import (EmbeddingGCN, EmbeddingGCN2, EmbeddingGCN_reg, EmbeddingKWGCN, TMGCNLayer, create_matrix_M,
                      func_MProduct, split_slices)
from .ops import Band, EdgePlan, SliceCSR
from .data import (compute_f1, compute_MAP_MRR, create_node_features, load_data, print_f1, save_mat,  # noqa: E402
                   split_data)

__all__ = ["EmbeddingGCN", "EmbeddingGCN2", "EmbeddingGCN_reg", "EmbeddingKWGCN", "TMGCNLayer", "create_matrix_M", "func_MProduct",
           "split_slices", "Band", "EdgePlan", "SliceCSR", "load_data", "save_mat", "create_node_features", "split_data",
           "compute_f1", "compute_MAP_MRR", "print_f1"]
__version__ = "0.1.0"
