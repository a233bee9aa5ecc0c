"""Host-side mirror of mufasa/spec_models: per-fittype constants and limits (meta_model.py), no pyspeckit."""
from .meta_model import MetaModel  # noqa: F401
