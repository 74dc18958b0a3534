"""``jax.numpy`` subset for traced rules: ``import jaxabm_b200.numpy as jnp`` in a model file that is
to run through the rule tracer (``jaxabm_b200/trace.py``).  Every function records the operation
on symbolic values; nothing here computes on the host."""
from .trace import numpy as _ns

float32, int32, bool_ = _ns.float32, _ns.int32, _ns.bool_
pi, e, inf, nan = _ns.pi, _ns.e, _ns.inf, _ns.nan
where, minimum, maximum, clip = _ns.where, _ns.minimum, _ns.maximum, _ns.clip
abs, absolute, sqrt, exp, log, log1p, tanh, power = (_ns.abs, _ns.absolute, _ns.sqrt, _ns.exp, _ns.log, _ns.log1p,
                                                     _ns.tanh, _ns.power)
logical_and, logical_or, logical_not = _ns.logical_and, _ns.logical_or, _ns.logical_not
sum, mean, max, min = _ns.sum, _ns.mean, _ns.max, _ns.min
asarray, array, nan_to_num = _ns.asarray, _ns.array, _ns.nan_to_num
