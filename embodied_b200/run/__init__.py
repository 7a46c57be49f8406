from .train import train
from .train_eval import train_eval
