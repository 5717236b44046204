"""B200-native (sm_100a) Hierarchical-LSTM teacher-student hot path of
shwetabhardwaj44/EfficientVideoClassification_Youtube8M behind the reference's plugin surface."""
__version__ = "0.1.0"
