"""stdlib logging front-end with the reference's call shape (rubix/logger.py:10-61):
``get_logger(config=None)`` where config may carry log_level / log_file_path / format."""

import logging

_DEFAULT = {"log_level": "WARNING", "log_file_path": None,
            "format": "%(asctime)s - %(name)s - %(levelname)s - %(message)s"}


def get_logger(config=None) -> logging.Logger:
    cfg = dict(_DEFAULT)
    if config:
        cfg.update({k: v for k, v in config.items() if v is not None or k == "log_file_path"})
    logger = logging.getLogger("rubix")
    logger.setLevel(getattr(logging, str(cfg["log_level"]).upper(), logging.WARNING))
    if not logger.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter(cfg["format"]))
        logger.addHandler(h)
        if cfg["log_file_path"]:
            fh = logging.FileHandler(cfg["log_file_path"])
            fh.setFormatter(logging.Formatter(cfg["format"]))
            logger.addHandler(fh)
    return logger
