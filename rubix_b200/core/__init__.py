"""Drop-in mirrors of the ``rubix.core`` factories for the particle -> datacube path."""
from .data import Galaxy, GasData, RubixData, StarsData, make_rubix_data, reshape_array, get_reshape_data  # noqa: F401
from .ifu import (get_calculate_datacube, get_calculate_spectra, get_doppler_shift_and_resampling,  # noqa: F401
                  get_scale_spectrum_by_mass)
from .lsf import get_convolve_lsf  # noqa: F401
from .psf import get_convolve_psf  # noqa: F401
from .ssp import get_lookup_interpolation, get_ssp  # noqa: F401
from .telescope import (get_filter_particles, get_spatial_bin_edges, get_spaxel_assignment,  # noqa: F401
                        get_telescope)
from .pipeline import RubixPipeline  # noqa: F401
from .rotation import get_galaxy_rotation  # noqa: F401
from .noise import get_apply_noise  # noqa: F401
from .dust import get_extinction  # noqa: F401
from .fits import load_fits, store_fits  # noqa: F401
