# The reference's snvc/models/__init__.py:1-2 holds only commented-out imports of the (unshipped)
# global model classes; the restated global hot path lives in snvc_b200.models.stereonet.
