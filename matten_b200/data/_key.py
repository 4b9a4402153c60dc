"""Keys of the graph dict that flows through the modules (same names as the reference,
src/matten/data/_key.py:14-49, so batches and state_dicts are interchangeable)."""
from typing import Dict, Final

import torch

Type = Dict[str, torch.Tensor]

POSITIONS: Final[str] = "pos"
NODE_ATTRS: Final[str] = "node_attrs"
NODE_FEATURES: Final[str] = "node_features"
EDGE_INDEX: Final[str] = "edge_index"
EDGE_CELL_SHIFT: Final[str] = "edge_cell_shift"
EDGE_VECTORS: Final[str] = "edge_vectors"
EDGE_LENGTH: Final[str] = "edge_lengths"
EDGE_ATTRS: Final[str] = "edge_attrs"
EDGE_EMBEDDING: Final[str] = "edge_embedding"
EDGE_MESSAGE: Final[str] = "edge_message"
CELL: Final[str] = "cell"
NUM_NEIGH: Final[str] = "num_neigh"
ATOMIC_NUMBERS: Final[str] = "atomic_numbers"
SPECIES_INDEX: Final[str] = "species_index"
PER_ATOM_ENERGY: Final[str] = "atomic_energy"
TOTAL_ENERGY: Final[str] = "total_energy"
BATCH: Final[str] = "batch"

# private: per-batch index bookkeeping built by matten_b200.graph.GraphCache
GRAPH_CACHE: Final[str] = "_mt_graph"
