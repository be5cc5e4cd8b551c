// empty stub: OpenFst is absent; the trainer includes this header but uses nothing from it
