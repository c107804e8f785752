"""``transform`` hook of Universe (reference layers/dyn_range_comp.py:28-37).

Every shipped config uses the identity transform (UniverseGAN hard-codes ``transform=None``,
universe_gan.py:91); the compressed-magnitude STFT transform is out of scope.
"""


class IdentityTransform:
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, x, inv=False):
        return x
