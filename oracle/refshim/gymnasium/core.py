class Env:
    metadata = {}

    def __init__(self, *a, **k):
        pass

    def __class_getitem__(cls, item):
        return cls


class Wrapper(Env):
    def __init__(self, env=None, *a, **k):
        self.env = env

    def __getattr__(self, name):
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)


class ActionWrapper(Wrapper):
    pass


class ObservationWrapper(Wrapper):
    pass


ActType = ObsType = WrapperObsType = WrapperActType = object
