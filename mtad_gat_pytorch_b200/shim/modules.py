"""Drop-in shim for the reference's `modules.py` import surface (mtad_gat.py:4-11)."""
from mtad_gat_pytorch_b200.modules import (ConvLayer, FeatureAttentionLayer, TemporalAttentionLayer,  # noqa: F401
                                           GRULayer, RNNDecoder, ReconstructionModel, Forecasting_Model)
