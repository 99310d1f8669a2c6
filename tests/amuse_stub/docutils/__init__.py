"""Stand-in for the docutils package, which this image does not have.

amuse.support.literature imports docutils at module import time to format its citation list
(/root/reference/src/amuse/support/literature.py:3,18); nothing on the ph4 force path uses it.  With this stub on
PYTHONPATH the unmodified AMUSE framework imports and runs (tests/test_gpu_amuse.py)."""
