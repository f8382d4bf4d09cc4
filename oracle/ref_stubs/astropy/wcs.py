class WCS:  # placeholder: the Cryo-NIRSP WCS branch is outside the hot path
    def __init__(self, *a, **k):
        raise NotImplementedError("astropy stub: WCS is not available in this container")
