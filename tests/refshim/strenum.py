"""TEST-ONLY stand-in for the `strenum` package (needed only to import the reference's cli.py:17)."""
from enum import Enum


class StrEnum(str, Enum):
    @staticmethod
    def _generate_next_value_(name, start, count, last_values):
        return name

    def __str__(self):
        return str(self.value)
