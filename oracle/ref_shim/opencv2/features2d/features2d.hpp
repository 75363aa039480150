// SHIM (test infrastructure only): common/src/flow_clusterer.cpp:13 includes this header and uses nothing from it.
