from .identifiersplitting import split_identifier_into_parts

__all__ = ["split_identifier_into_parts"]
