from .flip import Flip
from .pick_and_place import PickAndPlace
from .push import Push
from .reach import Reach
from .slide import Slide
from .stack import Stack

__all__ = ["Reach", "Push", "Slide", "PickAndPlace", "Stack", "Flip"]
