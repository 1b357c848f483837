"""Empty stand-in so the reference's module-level `import matplotlib.pyplot` succeeds (CL:8)."""
