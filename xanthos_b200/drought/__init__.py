from .drought_stats import DroughtStats  # noqa: F401
