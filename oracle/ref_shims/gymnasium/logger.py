def warn(*a, **k):
    pass
