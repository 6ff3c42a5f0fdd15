"""``ceacoest.modelling.symoptim`` names used by /root/reference/symfem.py:8-48."""
from oracle.engine import Model  # noqa: F401
