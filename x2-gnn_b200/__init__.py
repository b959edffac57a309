"""x2gnn_b200: B200-native SBFTransformerConv hot path."""
__version__ = "0.1.0"
