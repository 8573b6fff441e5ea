class RecordConstructorArgs:
    def __init__(self, *a, **k):
        pass
