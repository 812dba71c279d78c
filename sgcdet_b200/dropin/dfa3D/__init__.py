"""Drop-in for ``packages/3D-deformable-attention/DFA3D/dfa3D`` backed by libsgcdet_b200.so."""
