"""Argument validation shared by the host layer (limits as reference ``validate.py:6-33``:
int64 indices, so chains up to 63 spins)."""


def _int_like(x):
    try:
        ok = int(x) == x and x >= 0
    except (TypeError, ValueError):
        ok = False
    if not ok:
        raise ValueError(f'Value must be a nonnegative integer (got "{x!r}")')
    return int(x)


def L(x):
    x = _int_like(x)
    if x > 63:
        raise ValueError('Spin chain lengths greater than 63 not supported.')
    return x


def spin_index(x):
    x = _int_like(x)
    if x > 62:
        raise ValueError('Spin chain lengths greater than 63 not supported.')
    return x
