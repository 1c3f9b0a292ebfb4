"""hydrobayes_b200 -- B200-native DREAM sampler behind the HydroBayes API (dream / dream_parallel / R_stat).

The package is a thin host layer over the C-ABI library libhydrobayes_b200.so (CUDA, sm_100a).  No CPU fallback.
"""
from ._abi import HbConfig, HbError, HbTape, default_config, lib, lib_path  # noqa: F401
from .sampler import DensitySession, HydroModel, R_stat, Sampler, dream, dream_parallel, log_density, mixture, prior_sample, t_distribution  # noqa: F401
from .game import de_mc, game  # noqa: F401
from .single_chain import am, am_sd, rwm  # noqa: F401

__all__ = ["dream", "dream_parallel", "game", "de_mc", "rwm", "am", "am_sd", "R_stat", "HydroModel", "mixture", "t_distribution", "log_density", "DensitySession", "prior_sample", "Sampler", "HbConfig",
           "HbTape", "HbError", "default_config", "lib", "lib_path"]
