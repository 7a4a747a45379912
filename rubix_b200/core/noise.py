"""rubix/core/noise.py mirror: ``get_apply_noise(config)`` -> ``apply_noise(rubixdata)``."""

from __future__ import annotations

from typing import Callable

from ..logger import get_logger
from .data import RubixData

SUPPORTED_NOISE_DISTRIBUTIONS = ["normal", "uniform"]


def get_apply_noise(config: dict) -> Callable:
    """rubix/core/noise.py:15-78 (same validation order and messages)."""
    if "noise" not in config["telescope"]:
        raise ValueError("Noise information not provided in telescope config")
    if "signal_to_noise" not in config["telescope"]["noise"]:
        raise ValueError("Signal to noise information not provided in noise config")
    if "noise_distribution" not in config["telescope"]["noise"]:
        raise ValueError(
            f"Noise distribution not provided in noise config. Currently supported distributions are: {SUPPORTED_NOISE_DISTRIBUTIONS}"
        )
    signal_to_noise = config["telescope"]["noise"]["signal_to_noise"]
    noise_distribution = config["telescope"]["noise"]["noise_distribution"]
    logger = get_logger(config.get("logger", None))

    def apply_noise(rubixdata: RubixData) -> RubixData:
        from .. import ops
        logger.info(
            f"Applying noise to datacube with signal to noise ratio: {signal_to_noise} and noise distribution: {noise_distribution}"
        )
        rubixdata.stars.datacube = ops.apply_noise(rubixdata.stars.datacube, signal_to_noise, noise_distribution)
        return rubixdata

    return apply_noise
