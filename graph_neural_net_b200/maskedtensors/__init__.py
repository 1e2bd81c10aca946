from . import maskedtensor  # noqa: F401
