"""Oracle-side stand-in for the absent third-party ``ceacoest`` package.

Put ``oracle/shim`` on ``sys.path`` to import the reference's ``fem.py`` /
``symfem.py`` unchanged on top of ``oracle.engine`` (test infrastructure only).
"""
