"""Drop-in for the reference's `Data2Seq/` package (tokenizers feeding the shared encoder)."""
from . import Acoustic, Graph, Hyper_Spectrum, Image, Text, Time_Series, Video
from .Data2Seq import Data2Seq

__all__ = ["Data2Seq", "Image", "Video", "Acoustic", "Time_Series", "Hyper_Spectrum", "Text", "Graph"]
