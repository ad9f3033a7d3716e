"""`--model=planar`: the B200-native planar plugin (implementation in marf_b200/planar.py)."""
from marf_b200.planar import Graph, ImplicitMask, Model, NeuralImageFunction, PosEmbedding  # noqa: F401
