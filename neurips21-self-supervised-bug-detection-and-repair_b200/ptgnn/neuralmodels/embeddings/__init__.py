from .strelementrepresentationmodel import (
    CharUnitEmbedder,
    StrElementRepresentationModel,
    SubtokenUnitEmbedder,
    TokenUnitEmbedder,
)

__all__ = ["StrElementRepresentationModel", "SubtokenUnitEmbedder", "TokenUnitEmbedder", "CharUnitEmbedder"]
