"""B200-native engine for the Hass–Hertäg–Durstewitz PFC network (drop-in for the reference's Brian 2 simulation loop)."""
__version__ = "0.1.0"
