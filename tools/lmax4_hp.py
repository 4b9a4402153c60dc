"""Hyper-parameters of the lmax-4 architecture of the pretrained 20230627 model (pretrained/20230627/config_final.yaml:24-42)."""
HP = {
    "species_embedding_dim": 16, "irreps_edge_sh": "0e + 1o + 2e + 3o + 4e", "num_radial_basis": 8,
    "radial_basis_start": 0.0, "radial_basis_end": 5.0, "radial_basis_type": "bessel", "num_layers": 3,
    "invariant_layers": 2, "invariant_neurons": 32, "average_num_neighbors": 28.0,
    "conv_layer_irreps": "32x0o+32x0e + 16x1o+16x1e + 4x2o+4x2e + 2x3o+2x3e + 2x4e",
    "nonlinearity_type": "gate", "normalization": "batch", "resnet": True,
    "conv_to_output_hidden_irreps_out": "16x0e + 2x2e + 4e", "output_format": "irreps",
    "output_formula": "ijkl=jikl=klij", "reduce": "mean",
}
