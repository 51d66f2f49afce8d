def pyramid_expand(*a, **k):  # only used by the reference's visualisation code
    raise NotImplementedError
