"""component -- drop-in for pycuda-euler's ``pycomponent`` (src/eulercuda/pycomponent.py).

``find_component_device`` runs to its fix-point (the reference loop runs once): the
label of a node is the smallest node id of its component.  On the GPU that is a lock-free
union-find (hook larger root under smaller, atomicCAS) instead of the reference's eleven
Shiloach-Vishkin sub-steps; the sub-step wrappers are kept for callers and run the reference's
per-step semantics as individual sm_100a kernels.
"""
import logging

import numpy as np

import _native

module_logger = logging.getLogger('eulercuda.pycomponent')

SV_DTYPE = _native.SV_DTYPE


def find_component_device(d_v, d_D, length):
    """pycomponent.py:676 -- component label (minimum node id) per successor-graph node."""
    v = np.asarray(d_v, dtype=SV_DTYPE)[:int(length)]
    return _native.default_context().find_components(v)


def _c():
    return _native.default_context()


def component_step_init(d_v, d_D, d_Q, length):
    """pycomponent.py:17 -- D[t]=t, Q[t]=0."""
    return _c().compat_cc_step("init", int(length), v=d_v)


def component_step1_shortcutting_p1(d_v, d_prevD, d_D, d_Q, length, s):
    """pycomponent.py:68 -- curD[t] = prevD[prevD[t]]."""
    return _c().compat_cc_step("s1p1", int(length), prevD=d_prevD, D=d_D)


def component_step1_shortcutting_p2(d_v, d_prevD, d_D, d_Q, length, s):
    """pycomponent.py:128 -- Q[curD[t]] = s where the label changed."""
    return _c().compat_cc_step("s1p2", int(length), prevD=d_prevD, D=d_D, Q=d_Q, s=int(s))


def component_Step2_P1(d_v, d_prevD, d_D, d_Q, d_t1, d_val1, d_t2, d_val2, length, s):
    """pycomponent.py:189 -- propose hooks toward smaller neighbour labels."""
    return _c().compat_cc_step("s2p1", int(length), v=d_v, prevD=d_prevD, D=d_D, val1=d_val1, val2=d_val2)


def component_Step2_P2(d_v, d_prevD, d_D, d_Q, d_t1, d_val1, d_t2, d_val2, length, s):
    """pycomponent.py:280 -- atomicMin hooks, Q[val] = s."""
    return _c().compat_cc_step("s2p2", int(length), D=d_D, Q=d_Q, t1=d_t1, val1=d_val1, t2=d_t2, val2=d_val2, s=int(s))


def component_Step3_P1(d_v, d_prevD, d_D, d_Q, d_t1, d_val1, d_t2, d_val2, length, s):
    """pycomponent.py:373 -- stagnant trees propose hooks to any different neighbour label."""
    return _c().compat_cc_step("s3p1", int(length), v=d_v, D=d_D, Q=d_Q, val1=d_val1, val2=d_val2, s=int(s))


def component_Step3_P2(d_v, d_prevD, d_D, d_Q, d_t1, d_val1, d_t2, d_val2, length, s):
    """pycomponent.py:455 -- atomicMin hooks."""
    return _c().compat_cc_step("s3p2", int(length), D=d_D, t1=d_t1, val1=d_val1, t2=d_t2, val2=d_val2)


def component_step4_P1(d_v, d_D, d_val1, length):
    """pycomponent.py:535 -- val1[t] = curD[curD[t]]."""
    return _c().compat_cc_step("s4p1", int(length), D=d_D)


def component_step4_P2(d_v, d_D, d_val1, length):
    """pycomponent.py:583 -- curD[t] = val1[t]."""
    return _c().compat_cc_step("s4p2", int(length), D=d_D, val1=d_val1)


def component_step5(d_Q, length, d_sptemp, s):
    """pycomponent.py:633 -- 1 if any Q[t] == s (another round is needed)."""
    return _c().compat_cc_step("s5", int(length), Q=d_Q, s=int(s))
