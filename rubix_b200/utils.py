"""Config helpers with the reference's names (rubix/utils.py:10-66)."""

from typing import Dict, Union

import yaml

from .config import PIPELINES


def read_yaml(path_to_file: str) -> dict:
    try:
        with open(path_to_file, "r") as fh:
            return yaml.safe_load(fh)
    except Exception as e:  # same error type/message as rubix/utils.py:62-65
        raise RuntimeError(f"Something went wrong while reading yaml file {str(path_to_file)}") from e


def get_config(config: Union[str, Dict]) -> Dict:
    return read_yaml(config) if isinstance(config, str) else config


def get_pipeline_config(name: str):
    if name not in PIPELINES:
        raise ValueError(f"Pipeline {name} not found in the configuration")
    return PIPELINES[name]
