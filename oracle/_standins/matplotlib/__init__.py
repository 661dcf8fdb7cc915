"""Empty stand-in so that `import pygeon` succeeds without matplotlib.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference imports
matplotlib.pyplot at package import (src/pygeon/viz/plot_spanningtree.py:1);
nothing on the hot path touches it.
"""
