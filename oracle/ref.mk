# oracle/_ref: the reference's OWN discrete-BIC sources, compiled where they lie under /root/reference against the
# shim headers in oracle/shim/ (Boost is absent).  Only ref_driver.cpp is ours.  Output: oracle/_ref/libref_bic.so
# (git-ignored, travels to the GPU box).  Reference sources are never copied into the repo.
REF ?= /root/reference
CXX = $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
# the reference's own flags (urlearning/Jamroot:14-45): -std=c++11 -fno-strict-aliasing; -w: its warnings are not ours
CXXFLAGS = -std=c++11 -fno-strict-aliasing -O2 -fPIC -w -pthread -I shim -I $(REF)
SRC = base/bayesian_network.cpp base/skeleton.cpp ad_tree/ad_tree.cpp ad_tree/ad_node.cpp ad_tree/vary_node.cpp \
      scoring_function/log_likelihood_calculator.cpp scoring_function/bic_scoring_function.cpp scoring_function/score_calculator.cpp
OUT = _ref
all: $(OUT)/libref_bic.so
$(OUT)/libref_bic.so: ref_driver.cpp $(addprefix $(REF)/urlearning/,$(SRC)) $(wildcard shim/boost/*.hpp)
	@mkdir -p $(OUT)
	$(CXX) $(CXXFLAGS) -shared -Wl,-Bsymbolic -o $@ ref_driver.cpp $(addprefix $(REF)/urlearning/,$(SRC))
