"""Vector mirrors of the reference environments (one class per env family)."""
