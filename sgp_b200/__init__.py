"""sgp_b200 — B200-native (sm_100a) implementation of the SGP training-free spatiotemporal
encoder hot path, behind the reference's encoder / preprocess API."""
from .encoders import SGPEncoder, SGPSpatialEncoder, SGPTemporalEncoder
from .preprocessing import (ShiftOperator, preprocess_adj, preprocess_dataset,
                            reservoir_preprocessing_, sgp_spatial_embedding)
from .reservoir import Reservoir, ReservoirLayer
from .utils import encode_dataset, self_normalizing_activation

__all__ = ["SGPEncoder", "SGPSpatialEncoder", "SGPTemporalEncoder", "ShiftOperator", "Reservoir",
           "ReservoirLayer", "preprocess_adj", "preprocess_dataset", "reservoir_preprocessing_",
           "sgp_spatial_embedding", "encode_dataset", "self_normalizing_activation"]
