"""B200-native FEM assembly-and-solve hot path with PyGeoN's operator API."""
