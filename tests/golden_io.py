"""Helpers to load the golden vectors in tests/golden/ (made by make_golden.py)."""

from __future__ import annotations

import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")) as f:
        return {k: f[k] for k in f.files}


def catalog_arrays(g: dict, prefix: str) -> dict:
    """per-catalog arrays: ra, dec, patch, optional z / w / kappa, centers, radii"""
    out = dict(
        ra=g[f"{prefix}_ra"], dec=g[f"{prefix}_dec"], patch=g[f"{prefix}_patch"],
        z=g.get(f"{prefix}_z"), w=g.get(f"{prefix}_w"), kappa=g.get(f"{prefix}_kappa"),
        centers=g[f"{prefix}_centers"], radii=g[f"{prefix}_radii"],
    )
    return out


def config_of(g: dict) -> dict:
    rweight = float(g["rweight"])
    res = int(g["resolution"])
    return dict(
        zedges=g["zedges"], closed=str(g["closed"]), ang_min=g["ang_min"], ang_max=g["ang_max"],
        rweight=None if np.isnan(rweight) else rweight, resolution=None if res < 0 else res,
        rmin=g["rmin"], rmax=g["rmax"], zmin=float(g["zmin"]), zmax=float(g["zmax"]),
        max_angle=float(g["max_angle"][0]),
    )


def links_of(g: dict) -> dict[int, set[int]]:
    links: dict[int, set[int]] = {}
    for i, j in g["links"]:
        links.setdefault(int(i), set()).add(int(j))
    return links
