"""Mirror of the reference package path ``nn.tmGlow`` (tmglow/nn/tmGlow.py)."""
from .tmGlow import TMGlow  # noqa: F401
