"""abcdez.jl_b200 -- B200-native implementation of the ABCdeZ.jl particle hot path.

The directory name contains a dot, so import it through the alias module at the repo root:
``import abcdez_b200`` (abcdez_b200.py registers this package under that name).
"""
from .host import *  # noqa: F401,F403
from .host import (ABCdeZError, Context, Factored, Model, compile_model, Population, abcdesmc, abcdemc, abcdesmc_batch, lib, model_names,
                   wsample_stratified, default_context, shard_range, EXPORTS, LIB_PATH)
from . import host, build, dist  # noqa: F401
