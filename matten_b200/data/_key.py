"""Names of the entries of the graph dict that flows through the modules.

The strings are the reference's (src/matten/data/_key.py:14-49): a batch produced for the reference can be fed to these
modules and vice versa.  They are generated from one table (constant name -> dict key, grouped by what writes them)."""
from typing import Dict

import torch

Type = Dict[str, torch.Tensor]

_TABLE = {
    # written by the data pipeline (neighbour list + collate)
    "inputs": dict(POSITIONS="pos", CELL="cell", ATOMIC_NUMBERS="atomic_numbers", BATCH="batch",
                   EDGE_INDEX="edge_index", EDGE_CELL_SHIFT="edge_cell_shift", NUM_NEIGH="num_neigh"),
    # written by the embedding modules
    "embeddings": dict(SPECIES_INDEX="species_index", NODE_ATTRS="node_attrs", NODE_FEATURES="node_features",
                       EDGE_VECTORS="edge_vectors", EDGE_LENGTH="edge_lengths", EDGE_ATTRS="edge_attrs",
                       EDGE_EMBEDDING="edge_embedding"),
    # reserved by the reference, unused on the tensor-property path
    "reserved": dict(EDGE_MESSAGE="edge_message", PER_ATOM_ENERGY="atomic_energy", TOTAL_ENERGY="total_energy"),
    # private to this package: per-batch index bookkeeping built by matten_b200.graph.GraphCache
    "private": dict(GRAPH_CACHE="_mt_graph"),
}
for _group in _TABLE.values():
    globals().update(_group)
ALL_KEYS = {name: key for group in _TABLE.values() for name, key in group.items()}
