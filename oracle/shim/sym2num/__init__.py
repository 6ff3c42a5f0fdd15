"""Oracle-side stand-in for ``sym2num`` (only imported, never called directly:
/root/reference/attas_sp_ml.py:10)."""
