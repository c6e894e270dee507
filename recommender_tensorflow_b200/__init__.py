"""recommender-tensorflow_b200: B200-native DeepFM / wide&deep train step behind the reference's
model_fn / feature-column API.  See DESIGN.md."""
from . import feature_column  # noqa: F401

__all__ = ["feature_column"]
