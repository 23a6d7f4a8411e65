def cmp(x, y):
    """Python-2 cmp(); int() casts matter for numpy bools."""
    return int(x > y) - int(x < y)
