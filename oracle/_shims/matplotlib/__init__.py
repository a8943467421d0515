"""Stub: the reference imports matplotlib transitively (models_2020/transformer/encoder_layer.py:1-2)
but never plots on the inference path.  TEST INFRASTRUCTURE ONLY."""


def use(*a, **k):
    pass
