"""Import stand-in for matplotlib (utils/evaluation_utils.py:18, utils/vis_utils.py:6): the reference imports pyplot at module level but
only plots inside the offline mAP evaluation (out of scope).  Any attempt to draw raises."""


def use(*a, **k):
    pass
