class MeshFix:
    pass
