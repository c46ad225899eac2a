// refshim/stubs/matplotlibcpp.h — TEST INFRASTRUCTURE ONLY.  The reference's Trainer.cpp includes its vendored matplotlib
// bridge (lib/matplotlibcpp.h, needs Python.h + numpy); plotting is out of scope, so every call is accepted and ignored.
#pragma once
#include <map>
#include <string>
#include <vector>
namespace matplotlibcpp {
template <class... A> bool plot(const A&...) { return true; }
// plot(x, y, "r", {{"label", "model"}}): a braced keyword map cannot bind to a deduced pack
template <class X, class Y> bool plot(const X&, const Y&, const std::string&, const std::map<std::string, std::string>&) { return true; }
template <class... A> bool named_plot(const A&...) { return true; }
template <class... A> bool scatter(const A&...) { return true; }
template <class... A> bool quiver(const A&...) { return true; }
template <class... A> void figure(const A&...) {}
template <class... A> void subplot(const A&...) {}
template <class... A> void title(const A&...) {}
template <class... A> void xlabel(const A&...) {}
template <class... A> void ylabel(const A&...) {}
template <class... A> void legend(const A&...) {}
template <class... A> void axis(const A&...) {}
template <class... A> void xlim(const A&...) {}
template <class... A> void ylim(const A&...) {}
template <class... A> void grid(const A&...) {}
template <class... A> void pause(const A&...) {}
template <class... A> void save(const A&...) {}
template <class... A> void show(const A&...) {}
template <class... A> void clf(const A&...) {}
template <class... A> void close(const A&...) {}
template <class... A> void draw(const A&...) {}
}  // namespace matplotlibcpp
