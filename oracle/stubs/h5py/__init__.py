class File:
    def __init__(self, *a, **k):
        raise NotImplementedError("h5py stand-in")
