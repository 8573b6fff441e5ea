def __getattr__(name):
    def _nop(*a, **k):
        raise NotImplementedError("refshim matplotlib stub: %s" % name)
    return _nop
