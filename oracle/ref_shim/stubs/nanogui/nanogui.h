// Stub for nanogui -- TEST INFRASTRUCTURE. Forward declarations used by src/progressview.hpp.
#pragma once
#include <functional>
namespace nanogui
{
class Screen;
class Window;
class Label;
class ProgressBar;
class Widget;
class Popup;
} // namespace nanogui
