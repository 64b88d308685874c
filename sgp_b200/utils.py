"""``encode_dataset`` — the boundary caller of the reference (lib/utils.py:10-47) — and the
``self_normalizing_activation`` helper name it exports (lib/utils.py:50-51)."""
from __future__ import annotations

import logging
from time import time

import torch

logger = logging.getLogger("sgp_b200")


def _as_list(obj):
    if isinstance(obj, (str, bytes)) or not hasattr(obj, "__iter__"):
        return [obj]
    return list(obj)


def encode_dataset(dataset, encoder_class, encoder_kwargs, encode_exogenous=True, keep_raw=False,
                   save_path=None):
    """Encode ``dataset`` (a tsl ``SpatioTemporalDataset`` or any object with the same
    ``exogenous / get_tensors / edge_index / edge_weight / add_exogenous / set_input_map``
    surface) and register the result as exogenous ``encoded_x``.  Same glue and same quirk as the
    reference: a non-bool ``encode_exogenous`` raises NameError (lib/utils.py:19-22)."""
    if isinstance(encode_exogenous, bool):
        preprocess_exogenous = dataset.exogenous.keys() if encode_exogenous else []
    preprocess_exogenous = _as_list(preprocess_exogenous)  # noqa: F821 - NameError kept on purpose

    x, _ = dataset.get_tensors(['data'] + preprocess_exogenous, preprocess=True, cat_dim=-1)
    encoder = encoder_class(**encoder_kwargs)

    start = time()
    encoded_x = encoder(x, edge_index=dataset.edge_index, edge_weight=dataset.edge_weight)
    elapsed = int(time() - start)

    if save_path is not None:
        torch.save(encoded_x, save_path)
    logger.info(f"Dataset encoded in {elapsed // 60}:{elapsed % 60:02d} minutes.")

    dataset.add_exogenous('encoded_x', encoded_x, add_to_input_map=False)
    input_map = {'x': ['encoded_x']}
    u = ([] if encode_exogenous else ['u']) + (['data'] if keep_raw else [])
    if len(u):
        input_map['u'] = u
    dataset.set_input_map(input_map)
    return dataset


def self_normalizing_activation(x: torch.Tensor, r: float = 1.0):
    """Name kept for importers of lib.utils; inside the reservoir this runs in the scan kernel
    (activation code SGP_ACT_SELF_NORM), this torch expression is only for external callers."""
    return r * torch.nn.functional.normalize(x, p=2, dim=-1)
