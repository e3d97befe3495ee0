"""B200-native coupling-matrix assembly (exact intersections) + C / C^T SpMV."""
__version__ = "0.1.0"
