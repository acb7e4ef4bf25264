"""zett_b200: B200-native implementation of ZeTT's embedding-prediction hot path."""
