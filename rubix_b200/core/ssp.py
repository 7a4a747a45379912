"""rubix/core/ssp.py mirrors."""

from __future__ import annotations

from typing import Callable

from ..logger import get_logger
from ..ssp import SSPGrid, get_ssp_template


def get_ssp(config: dict) -> SSPGrid:
    """rubix/core/ssp.py:12-34 (same checks, same messages)."""
    if "ssp" not in config:
        raise ValueError("Configuration does not contain 'ssp' field")
    if "template" not in config["ssp"]:
        raise ValueError("Configuration does not contain 'template' field")
    if "name" not in config["ssp"]["template"]:
        raise ValueError("Configuration does not contain 'name' field")
    return get_ssp_template(config["ssp"]["template"]["name"])


def get_method(config: dict) -> str:
    """rubix/core/ssp.py:57-62: the default is *cubic* when ``ssp.method`` is absent."""
    logger = get_logger(config.get("logger", None))
    if "method" not in config["ssp"]:
        logger.debug("Method not defined, using default method: cubic")
        return "cubic"
    logger.debug(f"Using method defined in config: {config['ssp']['method']}")
    return config["ssp"]["method"]


def get_lookup_interpolation(config: dict) -> Callable:
    """rubix/core/ssp.py:38-65: ``lookup(metallicity, age) -> (n, L)`` on the GPU."""
    ssp = get_ssp(config)
    return ssp.get_lookup_interpolation(method=get_method(config))
