// empty stub: only forward declarations are needed by hmm/transition-model.h
