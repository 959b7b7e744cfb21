def configure(*a, **kw):
    return None


def warn(*a, **kw):
    return None
