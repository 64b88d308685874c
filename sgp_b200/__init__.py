"""sgp_b200 — B200-native (sm_100a) implementation of the SGP training-free spatiotemporal
encoder hot path, behind the reference's encoder / preprocess API."""
from .encoders import SGPEncoder, SGPSpatialEncoder, SGPTemporalEncoder
from .decoder import GroupedPointwiseConv
from .graph_reservoir import GESNEncoder, GESNLayer, GraphESN
from .preprocessing import (MeanOperator, OperatorChain, ShiftOperator, SparseAdj, preprocess_adj, preprocess_dataset,
                            reservoir_preprocessing_, sgp_collate_features, sgp_spatial_embedding,
                            sgp_spatial_support)
from .reservoir import Reservoir, ReservoirLayer
from .sampler import IIDSampler
from .utils import encode_dataset, self_normalizing_activation

__all__ = ["SGPEncoder", "SGPSpatialEncoder", "SGPTemporalEncoder", "ShiftOperator", "Reservoir",
           "ReservoirLayer", "preprocess_adj", "preprocess_dataset", "reservoir_preprocessing_",
           "sgp_spatial_embedding", "encode_dataset", "self_normalizing_activation",
           "sgp_spatial_support", "sgp_collate_features", "OperatorChain", "MeanOperator", "SparseAdj", "IIDSampler",
           "GroupedPointwiseConv", "GESNEncoder", "GESNLayer", "GraphESN"]
