"""reak_b200 — B200-native batched propagator for ReaK's KTE-chain forward-dynamics path.

Public surface: `reak_b200.kte` (mirror of the ReaK::kte modelling API), `reak_b200.presets`
(the benchmark chains) and `reak_b200.kte_batch_propagator` (batched
get_state_derivative / RK4 get_next_state on the GPU through libreak_b200.so).
"""
from . import kte, presets  # noqa: F401
from .propagator import kte_batch_propagator  # noqa: F401

__all__ = ["kte", "presets", "kte_batch_propagator"]
