"""CPU oracle for the colloc-fem hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is imported by the product package; see
``oracle/engine.py`` for the contract and the parity status.
"""
