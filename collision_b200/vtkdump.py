"""Debug dump of one detection pass as a VTK PolyData file -- the counterpart of the reference's
vtkplotVectorSurface (vtk.cpp:14-93, called from updateAverageVelocity under debugging("CollisionImpulse"),
dcollid.cpp:684-694): every element of hseList as a cell, coloured red when one of its points collected an
impulse in the pass (collsn_num > 0) and green otherwise, with the accumulated collsnImpulse as a point vector.

No VTK dependency: the file is written as ASCII XML (.vtp), which ParaView / VTK read directly.  Pure host
code over arrays the caller already has (positions, and the accumulators of CollisionSolver3d.accumulators()
between detect() and apply()); it never touches the device.
"""
from __future__ import annotations

import numpy as np


def point_order(tri_idx: np.ndarray, bond_idx: np.ndarray) -> np.ndarray:
    """Vertex ids in order of first appearance in hseList (triangles, then bonds): the numbering the reference
    assigns through POINT::indx (vtk.cpp:37-44)."""
    flat = np.concatenate([np.asarray(tri_idx, np.int64).reshape(-1), np.asarray(bond_idx, np.int64).reshape(-1)])
    _, first = np.unique(flat, return_index=True)
    return flat[np.sort(first)]


def _fmt(a, per_line, fmt):
    a = np.asarray(a).reshape(-1, per_line)
    return "\n".join(" ".join(fmt % v for v in row) for row in a)


def vtkplotVectorSurface(fname: str, x: np.ndarray, tri_idx: np.ndarray, bond_idx: np.ndarray, imp: np.ndarray,
                         cnt: np.ndarray) -> dict:
    """x (V,3) positions, tri_idx (T,3), bond_idx (B,2), imp (V,3) accumulated collsnImpulse, cnt (V,) collsn_num.
    Returns {"points": n, "cells": n} like the reference's two progress lines."""
    x = np.asarray(x, np.float64).reshape(-1, 3)
    tri_idx = np.asarray(tri_idx, np.int64).reshape(-1, 3)
    bond_idx = np.asarray(bond_idx, np.int64).reshape(-1, 2)
    imp = np.asarray(imp, np.float64).reshape(-1, 3)
    cnt = np.asarray(cnt).reshape(-1)
    order = point_order(tri_idx, bond_idx)
    new_id = np.full(x.shape[0], -1, np.int64)
    new_id[order] = np.arange(order.size)
    red, green = (255, 0, 0), (0, 255, 0)
    hit = cnt > 0
    tri_col = np.where(hit[tri_idx].any(axis=1)[:, None], red, green) if tri_idx.size else np.zeros((0, 3), int)
    bond_col = np.where(hit[bond_idx].any(axis=1)[:, None], red, green) if bond_idx.size else np.zeros((0, 3), int)
    T, B = tri_idx.shape[0], bond_idx.shape[0]
    # VTK stores cell data for lines before polygons
    colors = np.concatenate([bond_col, tri_col]) if (T + B) else np.zeros((0, 3), int)
    out = ['<?xml version="1.0"?>',
           '<VTKFile type="PolyData" version="0.1" byte_order="LittleEndian">', "<PolyData>",
           f'<Piece NumberOfPoints="{order.size}" NumberOfVerts="0" NumberOfLines="{B}" NumberOfStrips="0" NumberOfPolys="{T}">',
           '<PointData Vectors="CollsnImpulse">',
           '<DataArray type="Float32" Name="CollsnImpulse" NumberOfComponents="3" format="ascii">',
           _fmt(imp[order], 3, "%.9g") if order.size else "", "</DataArray>", "</PointData>",
           '<CellData Scalars="CollsnRegion">',
           '<DataArray type="UInt8" Name="CollsnRegion" NumberOfComponents="3" format="ascii">',
           _fmt(colors, 3, "%d") if colors.size else "", "</DataArray>", "</CellData>",
           "<Points>", '<DataArray type="Float32" NumberOfComponents="3" format="ascii">',
           _fmt(x[order], 3, "%.9g") if order.size else "", "</DataArray>", "</Points>"]
    for tag, idx, n in (("Lines", bond_idx, 2), ("Polys", tri_idx, 3)):
        out += [f"<{tag}>", '<DataArray type="Int64" Name="connectivity" format="ascii">',
                _fmt(new_id[idx], n, "%d") if idx.size else "", "</DataArray>",
                '<DataArray type="Int64" Name="offsets" format="ascii">',
                " ".join(str(n * (i + 1)) for i in range(idx.shape[0])), "</DataArray>", f"</{tag}>"]
    out += ["</Piece>", "</PolyData>", "</VTKFile>", ""]
    with open(fname, "w") as f:
        f.write("\n".join(out))
    return {"points": int(order.size), "cells": int(T + B)}
