class EllipticalCopula:
    pass
