"""Drop-in for flair_zonal_detection/model_utils.py (same names, arguments and results)."""
from __future__ import annotations

import logging
from copy import deepcopy
from typing import Any, Dict

from ..flair_hub.models.checkpoint import load_checkpoint
from ..flair_hub.models.flair_model import FLAIR_HUB_Model
from .raster import open_raster

logger = logging.getLogger(__name__)


def get_resolution(path) -> float:
    """model_utils.py:11-16."""
    return abs(open_raster(path).res[0])


def compute_patch_sizes(config: Dict[str, Any]) -> Dict[str, int]:
    """model_utils.py:19-35."""
    patch_sizes = {}
    target_res = config['reference_resolution']
    for mod, active in config['modalities']['inputs'].items():
        if not active:
            continue
        mod_res = get_resolution(config['modalities'][mod]['input_img_path'])
        scale = mod_res / target_res
        patch_sizes[mod] = int(round(config['img_pixels_detection'] / scale))
    logger.info('PATCH SIZES ---> %s', patch_sizes)
    return patch_sizes


def prepare_model_config(config: Dict[str, Any]) -> Dict[str, Any]:
    """model_utils.py:38-109: zonal config -> training-style model config."""
    cfg = deepcopy({k: v for k, v in config.items() if k != 'device'})
    if 'device' in config:
        cfg['device'] = config['device']
    cfg.setdefault('models', {})
    if 'monotemp_arch' in config:
        cfg['models']['monotemp_model'] = {'arch': config['monotemp_arch'], 'new_channels_init_mode': 'random'}
    if 'multitemp_model_ref_date' in config:
        cfg['models']['multitemp_model'] = {
            'ref_date': config['multitemp_model_ref_date'], 'encoder_widths': [64, 64, 64, 128],
            'decoder_widths': [32, 32, 64, 128], 'out_conv': [32, 19], 'str_conv_k': 3, 'str_conv_s': 1,
            'str_conv_p': 1, 'agg_mode': "att_group", 'encoder_norm': "group", 'n_head': 16, 'd_model': 256,
            'd_k': 4, 'pad_value': 0, 'padding_mode': "reflect"}
    cfg.setdefault("labels", [t["name"] for t in cfg["tasks"] if t.get("active", False)])
    cfg.setdefault("labels_configs", {
        t["name"]: {"value_name": list(t["class_names"].values())} for t in cfg["tasks"] if t.get("active", False)})
    cfg["modalities"].setdefault("inputs_channels", {
        mod: cfg["modalities"].get(mod, {}).get("channels", []) for mod in cfg["modalities"]["inputs"]})
    cfg["modalities"].setdefault("aux_loss", {mod: False for mod in cfg["modalities"]["inputs"]})
    dem_cfg = cfg["modalities"].get("DEM_ELEV", {})
    cfg["modalities"].setdefault("pre_processings", {
        "calc_elevation": dem_cfg.get("calc_elevation", False),
        "calc_elevation_stack_dsm": dem_cfg.get("calc_elevation_stack_dsm", False),
        "filter_sentinel2": False, "filter_sentinel2_max_cloud": 100, "filter_sentinel2_max_snow": 100,
        "filter_sentinel2_max_frac_cover": 1.0, "temporal_average_sentinel2": False,
        "temporal_average_sentinel1": False, "use_augmentation": False})
    cfg.setdefault("paths", {})["ckpt_model_path"] = config["model_weights"]
    return cfg


def build_inference_model(config: Dict[str, Any], patch_sizes: Dict[str, int]) -> FLAIR_HUB_Model:
    """model_utils.py:112-119: build, load ``config['model_weights']``, ``.eval()``; the caller
    moves it to the device (scripts/run_fast_aigle_segmentation.py:84)."""
    model_cfg = prepare_model_config(config)
    model = FLAIR_HUB_Model(config=model_cfg, img_input_sizes=patch_sizes,
                            max_batch=int(config.get('batch_size', 16)))
    load_checkpoint(model_cfg, model)
    return model.eval()
