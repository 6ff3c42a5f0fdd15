"""``ceacoest.optim`` names used by /root/reference/fem.py:6-62."""
from oracle.engine import Decision, Problem  # noqa: F401
