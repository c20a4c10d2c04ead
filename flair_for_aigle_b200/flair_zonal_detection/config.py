"""Zonal config loading: drop-in for flair_zonal_detection/config.py.

``load_config`` accepts the reference's YAML (configs/config_model_zonal_segmentation.yaml), the
same keys as JSON, or the outer product JSON (configs/config_aigle_aerial_segmentation.json)
whose ``model_config`` points at the YAML ("config JSON": accept both).
"""
from __future__ import annotations

import json
import logging
import os

import yaml

logger = logging.getLogger(__name__)

REQUIRED_KEYS = ['output_path', 'output_name', 'model_weights', 'img_pixels_detection',
                 'margin', 'modalities', 'tasks', 'output_px_meters']


def load_config(path: str) -> dict:
    """config.py:6-11 (+ JSON).  An outer product JSON is followed through ``model_config``; its
    other keys are kept under ``config['aigle']``."""
    with open(path, 'r') as f:
        text = f.read()
    if path.lower().endswith('.json'):
        cfg = json.loads(text)
        if 'model_config' in cfg and 'modalities' not in cfg:
            inner = load_config(cfg['model_config'])
            inner['aigle'] = {k: v for k, v in cfg.items() if k != 'model_config'}
            return inner
        return cfg
    return yaml.safe_load(text)


def validate_config(config: dict) -> None:
    """config.py:14-29: same required keys, same exceptions; creates ``output_path``."""
    for key in REQUIRED_KEYS:
        if key not in config:
            raise ValueError(f"Missing required config key: {key}")
    if not os.path.isfile(config['model_weights']):
        raise FileNotFoundError(f"Model weights not found at: {config['model_weights']}")
    os.makedirs(config['output_path'], exist_ok=True)


def config_recap_1(config: dict) -> None:
    used = ', '.join(m for m, a in config['modalities']['inputs'].items() if a)
    tasks = ', '.join(t['name'] for t in config['tasks'] if t['active'])
    logger.info("FLAIR-HUB ZONE DETECTION (B200) | output %s/%s.tif | modalities %s | tasks %s | type %s | "
                "weights %s | batch %s", config['output_path'], config['output_name'], used, tasks,
                config.get('output_type'), config['model_weights'], config.get('batch_size'))


def config_recap_2(config: dict) -> None:
    shape = config.get('image_shape_px', {})
    logger.info("image %sx%s px | ref res %s m/px | out res %s m/px | patch %s px margin %s px | resolutions %s",
                shape.get('height'), shape.get('width'), config['reference_resolution'], config['output_px_meters'],
                config['img_pixels_detection'], config['margin'], config.get('modality_resolutions'))
