"""prodsearch_b200: B200-native (sm_100a) implementation of the embedding-scoring hot path of
kepingbi/ProdSearch behind the reference's own nn.Module surface.  See DESIGN.md."""
from . import _lib  # noqa: F401

__all__ = ["ItemTransformerRanker", "ProdSearchModel", "ProductRanker", "ParagraphVector",
           "ParagraphVectorCorruption", "FSEncoder", "AVGEncoder", "get_vector_mean"]


def __getattr__(name):
    if name in ("ItemTransformerRanker", "ProdSearchModel"):
        from . import item_transformer as m
    elif name == "ProductRanker":
        from . import ps_model as m
    elif name == "ParagraphVector":
        from . import pv as m
    elif name == "ParagraphVectorCorruption":
        from . import pvc as m
    elif name in ("FSEncoder", "AVGEncoder", "get_vector_mean"):
        from . import text_encoder as m
    else:
        raise AttributeError(name)
    return getattr(m, name)
