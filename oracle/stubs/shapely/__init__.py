class Point:
    pass


class Polygon:
    pass
