"""rubix/core/rotation.py mirror: ``get_galaxy_rotation(config)`` -> ``rotate_galaxy(rubixdata)``."""

from __future__ import annotations

from typing import Callable

from ..logger import get_logger
from .data import RubixData


def get_galaxy_rotation(config: dict) -> Callable:
    """rubix/core/rotation.py:11-115 (same validation order and messages)."""
    if "rotation" not in config["galaxy"]:
        raise ValueError("Rotation information not provided in galaxy config")
    logger = get_logger(config.get("logger", None))
    rot = config["galaxy"]["rotation"]
    if "type" in rot:
        if rot["type"] not in ["face-on", "edge-on"]:
            raise ValueError("Invalid type provided in rotation information")
        alpha, beta, gamma = (0.0, 0.0, 0.0) if rot["type"] == "face-on" else (90.0, 0.0, 0.0)
    else:
        for key in ["alpha", "beta", "gamma"]:
            if key not in rot:
                raise ValueError(f"{key} not provided in rotation information")
        alpha, beta, gamma = rot["alpha"], rot["beta"], rot["gamma"]
    particle_types = config.get("data", {}).get("args", {}).get("particle_type", ["stars"])

    def rotate_galaxy(rubixdata: RubixData) -> RubixData:
        from .. import ops
        logger.info(f"Rotating galaxy with alpha={alpha}, beta={beta}, gamma={gamma}")
        for particle_type in ["stars", "gas"]:
            if particle_type in particle_types:
                component = getattr(rubixdata, particle_type)
                if component is None or component.coords is None:
                    continue
                assert component.velocity is not None, f"Velocities not found for {particle_type}. "
                assert component.mass is not None, f"Masses not found for {particle_type}. "
                # the reference uses the STELLAR half-mass radius for both components (rotation.py:88)
                # a galaxy sharded over ranks (b200.distributed) gets ONE rotation: the inertia sums are all-reduced
                comm = None
                if isinstance(config.get("b200"), dict) and config["b200"].get("distributed"):
                    from ..parallel import get_comm
                    comm = get_comm()
                coords, velocity, _ = ops.rotate_galaxy(component.coords, component.velocity, component.mass,
                                                        float(rubixdata.galaxy.halfmassrad_stars), alpha, beta, gamma,
                                                        comm=comm)
                component.coords, component.velocity = coords, velocity
        return rubixdata

    return rotate_galaxy
