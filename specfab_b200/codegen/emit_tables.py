"""Build-time: emit the sparse operator tables the operator-export kernels consume (placeholder)."""


def emit():
    return "// GENERATED tables (none yet)\n"
