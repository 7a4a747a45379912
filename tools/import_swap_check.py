"""The import swap of INTEGRATION.md section 1, executed: the mirror's factories (rubix_b200.core) are handed to the
REFERENCE'S OWN pipeline machinery -- rubix/pipeline/{transformer,abstract_pipeline,linear_pipeline}.py run from source
through tools/refshim.py (jit is the identity there) with the reference's rubix/config/pipeline_config.yml -- exactly as
rubix/core/pipeline.py:105-157 does with its own factories: registered by __name__, deep-copied by bound_transformer,
ordered by depends_on, composed into one expression.  Build container only (needs /root/reference).

Without a GPU the composed expression must stop at the first stage with the library's "no CPU fallback" error; with one
(``--run``) it is run on the TNG50 subset and the cube is compared with the mirror's own RubixPipeline.
"""

import copy
import os
import sys

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refshim  # noqa: E402

CONFIG = {
    "pipeline": {"name": "calc_ifu"},
    "logger": {"log_level": "WARNING", "log_file_path": None,
               "format": "%(asctime)s - %(name)s - %(levelname)s - %(message)s"},
    "telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6}, "lsf": {"sigma": 0.5},
                  "noise": {"signal_to_noise": 50.0, "noise_distribution": "normal"}},
    "cosmology": {"name": "PLANCK15"},
    "galaxy": {"dist_z": 0.1, "rotation": {"alpha": 20.0, "beta": -35.0, "gamma": 70.0}},
    "ssp": {"template": {"name": "BruzualCharlot2003"},
            "dust": {"extinction_model": "Cardelli89", "Rv": 3.1, "dust_grain_density": 3.5}},
    "data": {"args": {"particle_type": ["stars"]}},
}


def reference_pipeline_class():
    refshim.install()
    sys.modules.setdefault("rubix.pipeline", type(sys)("rubix.pipeline")).__path__ = []
    sys.modules["jax"].make_jaxpr = lambda f, **k: f
    refshim.load("rubix/pipeline/transformer.py")
    sys.modules["rubix.pipeline"].abstract_pipeline = refshim.load("rubix/pipeline/abstract_pipeline.py")
    return refshim.load("rubix/pipeline/linear_pipeline.py").LinearTransformerPipeline


def mirror_functions(cfg):
    from rubix_b200 import core
    # the order of rubix/core/pipeline.py:120-133
    return [core.get_galaxy_rotation(cfg), core.get_filter_particles(cfg), core.get_spaxel_assignment(cfg),
            core.get_calculate_spectra(cfg), core.get_reshape_data(cfg), core.get_scale_spectrum_by_mass(cfg),
            core.get_doppler_shift_and_resampling(cfg), core.get_extinction(cfg), core.get_calculate_datacube(cfg),
            core.get_convolve_psf(cfg), core.get_convolve_lsf(cfg), core.get_apply_noise(cfg)]


def main():
    from rubix_b200 import core
    Pipeline = reference_pipeline_class()
    cfgs = yaml.safe_load(open(os.path.join(refshim.REF, "rubix", "config", "pipeline_config.yml")))
    d = np.load(os.path.join(ROOT, "tests", "golden", "tng50_subset.npz"))
    for name in ("calc_ifu", "calc_dusty_ifu"):
        cfg = copy.deepcopy(CONFIG)
        cfg["pipeline"]["name"] = name
        pipe = Pipeline(cfgs[name], mirror_functions(cfg))
        func = pipe.compile_expression()                 # rubix/core/pipeline.py:160
        print(name, "assembled:", " -> ".join(pipe._names))
        rd = core.make_rubix_data(**{k: d[k] for k in d.files}, device=False)
        rd.galaxy.halfmassrad_stars = 2.5
        if "--run" in sys.argv and name == "calc_ifu":
            out = func(rd)
            own = core.RubixPipeline(cfg, data=core.make_rubix_data(**{k: d[k] for k in d.files}, device=False))
            own.data.galaxy.halfmassrad_stars = 2.5
            ref = own.run()
            a, b = out.stars.datacube.cpu().numpy(), ref.stars.datacube.cpu().numpy()
            print("ran through the reference machinery: cube", a.shape, "identical to RubixPipeline:", np.array_equal(a, b))
        else:
            try:
                func(rd)
                print(name, "ran (a CUDA device is present)")
            except RuntimeError as e:
                print(name, "stopped at the first stage:", e)
    return 0


if __name__ == "__main__":
    sys.exit(main())
