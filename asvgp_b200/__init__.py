"""asvgp_b200 — B200-native implementation of the ASVGP hot path behind the reference's own API
(`basis.B{1..6}Spline`, `inducing_features.SplineFeatures1D`, `gpr.GPR_1d`, `gpr.GPR_kron`).

    import asvgp_b200.basis as basis
    from asvgp_b200.gpr import GPR_1d

mirrors `import asvgp.basis as basis; from asvgp.gpr import GPR_1d` of the reference."""
__version__ = "0.1.0"
