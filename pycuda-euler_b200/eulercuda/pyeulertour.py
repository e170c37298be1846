"""eulertour -- drop-in for pycuda-euler's ``pyeulertour`` (src/eulercuda/pyeulertour.py)."""
import logging

import numpy as np

import _native
from .pyencode import getOptimalLaunchConfiguration  # noqa: F401  (re-exported like the reference, :10)
from .pycomponent import find_component_device

module_logger = logging.getLogger('eulercuda.pyeulertour')

EV_DTYPE, EE_DTYPE, SV_DTYPE, CE_DTYPE = _native.EV_DTYPE, _native.EE_DTYPE, _native.SV_DTYPE, _native.CE_DTYPE


def _c():
    return _native.default_context()


def assign_successor_device(d_ev, d_l, d_e, vcount, d_ee, ecount):
    """pyeulertour.py:18 -- pair the i-th entering with the i-th leaving edge of every vertex."""
    ee = _c().assign_successor(np.asarray(d_ev, dtype=EV_DTYPE)[:int(vcount)], d_l, d_e,
                               np.asarray(d_ee, dtype=EE_DTYPE)[:int(ecount)])
    return d_ev, ee


def construct_successor_graphP1_device(d_ee, d_v, ecount):
    """pyeulertour.py:110 -- v[t] = {vid=eid, n1=s, n2=ecount}."""
    ee = np.asarray(d_ee, dtype=EE_DTYPE)[:int(ecount)]
    v = np.zeros(ee.size, SV_DTYPE)
    v['vid'] = ee['eid'].astype(np.uint32)
    v['n1'] = ee['s']
    v['n2'] = int(ecount)
    return v


def construct_successor_graphP2_device(d_ee, d_v, ecount):
    """pyeulertour.py:165 -- predecessor links v[v[t].n1].n2 = v[t].vid (P1+P2 fused on device)."""
    return _c().successor_graph(np.asarray(d_ee, dtype=EE_DTYPE)[:int(ecount)])


def calculate_circuit_graph_vertex_data_device(d_D, d_C, length):
    """pyeulertour.py:220 -- C[D[t]] = 1."""
    D = np.asarray(d_D, dtype=np.uint32)[:int(length)]
    Cm, _, _, _ = _c().circuit_vertices(D)
    return D, Cm


def construct_circuit_Graph_vertex(d_C, d_cg_offset, ecount, d_cv):
    """pyeulertour.py:269 -- cv[offset[t]] = t where C[t] != 0."""
    Cm = np.asarray(d_C, dtype=np.uint32)[:int(ecount)]
    return np.flatnonzero(Cm).astype(np.uint32)


def calculate_circuit_graph_edge_data(d_ev, d_e, vcount, d_D, d_cg_offset, ecount, d_cedgeCount):
    """pyeulertour.py:308 -- circuit-edge count per smaller circuit id."""
    cg = _c().circuit_edges(np.asarray(d_ev, dtype=EV_DTYPE)[:int(vcount)], d_e, d_D, d_cg_offset)
    out = np.zeros(int(ecount), np.uint32)
    if cg.size:
        np.add.at(out, cg['c1'], 1)
    return out


def assign_circuit_graph_edge_data(d_ev, d_e, vcount, d_D, d_cg_offset, ecount, d_cg_edge_start, d_cedgeCount,
                                   cvCount, d_cg_edge, cecount):
    """pyeulertour.py:394 -- CircuitEdge records (returned already sorted by (c1, c2), :792)."""
    return _c().circuit_edges(np.asarray(d_ev, dtype=EV_DTYPE)[:int(vcount)], d_e, d_D, d_cg_offset)


def mark_spanning_euler_edges(d_ee, d_mark, ecount, d_cg_edge, cg_edgeCount, d_tree, treeCount):
    """pyeulertour.py:587 -- mark[min(e1, e2)] = 1 for every spanning-forest circuit edge."""
    cg = np.asarray(d_cg_edge, dtype=CE_DTYPE)[:int(cg_edgeCount)]
    tree = np.asarray(d_tree, dtype=np.uint32).ravel()[:int(treeCount)]
    return _c().mark_spanning(cg, tree, int(ecount))


def execute_swipe(d_ev, d_e, vcount, d_ee, d_mark, ecount):
    """pyeulertour.py:496 -- rotate successors across runs of marked entering edges."""
    ee = _c().swipe(np.asarray(d_ev, dtype=EV_DTYPE)[:int(vcount)], d_e,
                    np.asarray(d_ee, dtype=EE_DTYPE)[:int(ecount)], d_mark)
    return ee, d_mark


def executeSwipeDevice(d_ev, d_e, vcount, d_ee, ecount, d_cg_edge, cg_edgeCount, d_tree, treeCount):
    """pyeulertour.py:657"""
    d_mark = mark_spanning_euler_edges(d_ee, None, ecount, d_cg_edge, cg_edgeCount, d_tree, treeCount)
    d_ee, _ = execute_swipe(d_ev, d_e, vcount, d_ee, d_mark, ecount)
    return d_ee


def identify_contig_start(d_ee, d_contigStart, ecount):
    """pyeulertour.py:668 -- 1 for edges that are nobody's successor."""
    return _c().contig_starts(np.asarray(d_ee, dtype=EE_DTYPE)[:int(ecount)])


def findEulerDevice(d_ev, d_l, d_e, vcount, d_ee, ecount, d_cg_edge, cg_edgeCount, cg_vertexCount):
    """pyeulertour.py:715 -- returns (d_cg_edge, circuitGraphEdgeCount, cg_vertexCount).

    The reference mutates d_ee through its wrappers' return values but drops them (it rebinds
    locals only); callers that need the successor table call ``assign_successor_device``.  With
    <= 1 circuit it returns an empty edge array and 0 (B11) instead of raising NameError."""
    ev = np.asarray(d_ev, dtype=EV_DTYPE)[:int(vcount)]
    _, ee = assign_successor_device(ev, d_l, d_e, vcount, d_ee, ecount)
    v = construct_successor_graphP2_device(ee, None, ecount)
    D = find_component_device(v, None, int(ecount))
    _, cg_offset, _, ncirc = _c().circuit_vertices(D)
    cg = np.zeros(0, CE_DTYPE)
    if ncirc > 1:
        cg = _c().circuit_edges(ev, d_e, D, cg_offset)
    return cg, int(cg.size), int(ncirc)
