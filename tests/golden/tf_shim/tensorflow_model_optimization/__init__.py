"""Stub of tensorflow_model_optimization: only the PrunableLayer mix-in name that
nif/layers/{siren,mlp}.py inherit from.  Test infrastructure only."""
import types

sparsity = types.SimpleNamespace(keras=types.SimpleNamespace(PrunableLayer=type("PrunableLayer", (), {})))
