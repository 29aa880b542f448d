"""Synthetic idealised geometries for the BASELINE configs.

Initial fields are evaluated at mesh vertices with numpy (host side; the reference does this in
``src/reference_fields_module.f90:425-566`` on a square grid and then maps to the mesh).
"""
from __future__ import annotations

import numpy as np

SEC_PER_YEAR = 31556943.36


def halfar_H(H0, R0, x, y, t):
    """Halfar (1981) similarity solution as coded in ``src/reference_fields_module.f90:707-745``."""
    A_flow, rho, g = 1e-16, 910.0, 9.81
    Gamma = (2.0 / 5.0) * (A_flow / SEC_PER_YEAR) * (rho * g) ** 3.0
    t0 = 1.0 / (18.0 * Gamma) * (7.0 / 4.0) ** 3.0 * (R0**4.0) / (H0**7.0)
    tp = t * SEC_PER_YEAR + t0
    r = np.sqrt(x**2 + y**2)
    f1 = (t0 / tp) ** (1.0 / 9.0)
    f2 = (t0 / tp) ** (1.0 / 18.0)
    return H0 * f1 * np.maximum(0.0, 1.0 - (f2 * r / R0) ** (4.0 / 3.0)) ** (3.0 / 7.0)


def bueler_H(H0, R0, lam, x, y, t):
    """Bueler (2005) solution as coded in ``src/reference_fields_module.f90:747-793``."""
    A_flow, rho, g, n = 1e-16, 910.0, 9.81, 3.0
    alpha = (2.0 - (n + 1.0) * lam) / (5.0 * n + 3.0)
    beta = (1.0 + (2.0 * n + 1.0) * lam) / (5.0 * n + 3.0)
    Gamma = 2.0 / 5.0 * (A_flow / SEC_PER_YEAR) * (rho * g) ** n
    f1 = (2.0 * n + 1.0) / (n + 1.0)
    f2 = R0 ** (n + 1.0) / H0 ** (2.0 * n + 1.0)
    t0 = (beta / Gamma) * f1**n * f2
    tp = t * SEC_PER_YEAR
    r = np.sqrt(x**2 + y**2) / R0
    f4 = np.maximum(0.0, 1.0 - ((tp / t0) ** (-beta) * r) ** ((n + 1.0) / n))
    return H0 * (tp / t0) ** (-alpha) * f4 ** (n / (2.0 * n + 1.0))


def state_halfar(mesh, H0=5000.0, R0=300000.0, t=0.0):
    """BASELINE config 2: Halfar dome, Hb = 0, SL = -10000 (src/reference_fields_module.f90:462-465,521-533)."""
    x, y = mesh.V[:, 0], mesh.V[:, 1]
    return dict(benchmark="Halfar", Hi=halfar_H(H0, R0, x, y, t), Hb=np.zeros(mesh.nV), SL=np.full(mesh.nV, -10000.0),
                SMB_year=np.zeros(mesh.nV), BMB=np.zeros(mesh.nV))


def state_eismint1(mesh):
    """BASELINE config 1: EISMINT-1 moving margin, ice-free start (src/SMB_module.f90:188-235)."""
    r = np.hypot(mesh.V[:, 0], mesh.V[:, 1])
    return dict(benchmark="EISMINT_1", Hi=np.zeros(mesh.nV), Hb=np.zeros(mesh.nV), SL=np.full(mesh.nV, -10000.0),
                SMB_year=np.minimum(0.5, 1e-5 * (450000.0 - r)), BMB=np.zeros(mesh.nV))


def state_ssa_icestream(mesh, scale=1.0, Hb=-500.0, H_shelf=300.0):
    """BASELINE config 3 (synthetic; the reference has no runnable SSA-only benchmark):
    flat bed at ``Hb``, SL = 0; 1000 m grounded ice for r < r1, linear taper to an ``H_shelf`` thick shelf at
    r2, 0.1 m thin-shelf convention beyond.  ``scale`` shrinks the radii with the domain.  Physics switches
    follow the reference's only SSA benchmark, MISMIP_mod (flow factor, SMB 0.3 m/yr)."""
    r = np.hypot(mesh.V[:, 0], mesh.V[:, 1])
    r1, r2 = 1000e3 * scale, 1400e3 * scale
    Hi = np.where(r < r1, 1000.0, np.where(r < r2, 1000.0 - (1000.0 - H_shelf) * (r - r1) / (r2 - r1), 0.1))
    Hi[mesh.edge_index > 0] = 0.0
    return dict(benchmark="MISMIP_mod", Hi=Hi, Hb=np.full(mesh.nV, float(Hb)), SL=np.zeros(mesh.nV),
                SMB_year=np.full(mesh.nV, 0.3), BMB=np.zeros(mesh.nV))


# bench.py workload (BASELINE configs[2]): bed at -250 m so that the grounding line sits at ~280 m thick ice and the
# Tsai et al. grounding-line flux (use_analytical_GL_flux, as in config_MISMIP_mod_*) gives O(1 km/yr) velocities
CONFIG3 = dict(half_width=1800e3, Hb=-250.0, H_shelf=150.0, use_analytical_GL_flux=1)


def state_mismip(mesh, Hi0=100.0, half_width=750e3):
    """BASELINE config 4: MISMIP_mod sloping bed Hb = 720 - 778.5 r / 750 km, Hi = 100 m
    (src/reference_fields_module.f90:549-564)."""
    r = np.hypot(mesh.V[:, 0], mesh.V[:, 1])
    Hi = np.full(mesh.nV, Hi0)
    Hi[mesh.edge_index > 0] = 0.0
    return dict(benchmark="MISMIP_mod", Hi=Hi, Hb=720.0 - 778.5 * r / 750e3, SL=np.zeros(mesh.nV),
                SMB_year=np.full(mesh.nV, 0.3), BMB=np.zeros(mesh.nV))


def state_thermo_dome(mesh, benchmark="EISMINT_1", H0=3000.0, R0=500e3, nZ=15, zeta=None):
    """Inputs for update_ice_temperature on a land-based dome: Halfar-shaped ice, SMB of the EISMINT-1
    moving-margin experiment, a seasonal 2 m temperature around 270 K - 0.01 K/m * Hs (the EISMINT-1 surface temperature,
    ``src/climate_module.f90``), a uniform geothermal heat flux of 0.0545 W m^-2 and a temperature field that rises linearly
    from the surface value to 2 K below pressure melting at the bed.  The climate fields are host inputs in the reference."""
    x, y = mesh.V[:, 0], mesh.V[:, 1]
    r = np.hypot(x, y)
    Hi = H0 * np.maximum(0.0, 1.0 - (r / R0) ** (4.0 / 3.0)) ** (3.0 / 7.0)
    Hi[mesh.edge_index > 0] = 0.0
    if zeta is None:
        zeta = np.array([0.00, 0.10, 0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.90, 0.925, 0.95, 0.975, 0.99, 1.00])
    Ts = 270.0 - 0.01 * Hi
    month = np.arange(12)
    T2m = np.asfortranarray(Ts[:, None] + 8.0 * np.cos(2.0 * np.pi * (month[None, :] + 0.5) / 12.0))
    Tbed = 273.16 - 8.7e-4 * Hi - 2.0
    Ti = np.asfortranarray(Ts[:, None] + zeta[None, :] * (Tbed - Ts)[:, None])
    return dict(benchmark=benchmark, Hi=Hi, Hb=np.zeros(mesh.nV), SL=np.full(mesh.nV, -10000.0),
                SMB_year=np.minimum(0.5, 1e-5 * (450000.0 - r)), BMB=np.zeros(mesh.nV),
                T2m=T2m, GHF=np.full(mesh.nV, 0.0545 * SEC_PER_YEAR), Ti=Ti)
