"""Application-specific tools (reference ampis/applications)."""
from . import powder

__all__ = ['powder']
