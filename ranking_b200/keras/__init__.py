"""`tfr.keras` surface: losses, metrics, layers, model (scorer), utils."""
from ranking_b200.keras import layers
from ranking_b200.keras import losses
from ranking_b200.keras import metrics
from ranking_b200.keras import model
from ranking_b200.keras import utils
