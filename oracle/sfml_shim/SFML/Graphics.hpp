// No-op SFML 2.5 surface: just enough for the reference nocturne_core to compile headless (oracle build only).
// Nothing on the rollout path calls into it at run time. Test infrastructure, not product code.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
namespace sf {
using Uint8 = unsigned char;
template <class T> struct Vector2 { T x{}, y{}; Vector2() = default; Vector2(T a, T b) : x(a), y(b) {} };
using Vector2f = Vector2<float>; using Vector2u = Vector2<unsigned>; using Vector2i = Vector2<int>;
struct Color { Uint8 r{}, g{}, b{}, a{255}; Color() = default; Color(int R,int G,int B,int A=255):r(R),g(G),b(B),a(A){}
  static const Color Black, White, Red, Green, Blue, Yellow, Magenta, Cyan, Transparent; };
inline const Color Color::Black(0,0,0), Color::White(255,255,255), Color::Red(255,0,0), Color::Green(0,255,0),
  Color::Blue(0,0,255), Color::Yellow(255,255,0), Color::Magenta(255,0,255), Color::Cyan(0,255,255), Color::Transparent(0,0,0,0);
template <class T> struct Rect { T left{}, top{}, width{}, height{}; Rect() = default; Rect(T l,T t,T w,T h):left(l),top(t),width(w),height(h){} };
using FloatRect = Rect<float>;
struct Transform { Transform& scale(float,float){return *this;} Transform& rotate(float){return *this;} };
struct RenderStates { RenderStates() = default; RenderStates(const Transform&) {} };
enum PrimitiveType { Points, Lines, LineStrip, Triangles, TriangleStrip, TriangleFan, Quads };
struct Vertex { Vector2f position; Color color; Vertex() = default; Vertex(const Vector2f& p, const Color& c = Color()) : position(p), color(c) {} };
struct View { View() = default; View(const FloatRect&) {} View(const Vector2f&, const Vector2f&) {}
  void setRotation(float) {} void setViewport(const FloatRect&) {} };
class RenderTarget; 
class Drawable { public: virtual ~Drawable() = default; protected: friend class RenderTarget;
  virtual void draw(RenderTarget& target, RenderStates states) const = 0; };
class RenderTarget { public: virtual ~RenderTarget() = default;
  void draw(const Drawable&, const RenderStates& = RenderStates()) {}
  void draw(const Vertex*, std::size_t, PrimitiveType, const RenderStates& = RenderStates()) {}
  void setView(const View&) {} void clear(const Color& = Color()) {} virtual Vector2u getSize() const { return {1,1}; } };
struct Shape : Drawable { void setFillColor(const Color&) {} void setOutlineColor(const Color&) {} void setOutlineThickness(float) {}
  void setOrigin(float,float) {} void setPosition(float,float) {} void setPosition(const Vector2f&) {} void setRotation(float) {}
  protected: void draw(RenderTarget&, RenderStates) const override {} };
struct CircleShape : Shape { explicit CircleShape(float=0, std::size_t=30) {} };
struct RectangleShape : Shape { explicit RectangleShape(const Vector2f& = Vector2f()) {} };
struct ConvexShape : Shape { explicit ConvexShape(std::size_t=0) {} void setPointCount(std::size_t) {} void setPoint(std::size_t, const Vector2f&) {} };
struct VertexArray : Drawable { VertexArray() = default; explicit VertexArray(PrimitiveType, std::size_t=0) {} void append(const Vertex&) {}
  protected: void draw(RenderTarget&, RenderStates) const override {} };
struct ContextSettings { unsigned antialiasingLevel{}; };
struct Image { const Uint8* getPixelsPtr() const { return nullptr; } bool saveToFile(const std::string&) const { return false; } };
struct Texture { bool create(unsigned,unsigned){return true;} template<class W> void update(const W&) {} Image copyToImage() const { return {}; } };
class RenderTexture : public RenderTarget { public: bool create(unsigned,unsigned,const ContextSettings& = ContextSettings()){return true;}
  static unsigned getMaximumAntialiasingLevel(){return 0;} void display() {} const Texture& getTexture() const { return tex_; } private: Texture tex_; };
struct Font { bool loadFromFile(const std::string&) { return false; } };
struct Text : Drawable { Text(const std::string&, const Font&, unsigned=30) {} void setPosition(float,float) {} void setFillColor(const Color&) {}
  protected: void draw(RenderTarget&, RenderStates) const override {} };
struct Time { float asSeconds() const { return 1.f; } };
struct Clock { Time restart() { return {}; } };
struct VideoMode { VideoMode(unsigned,unsigned,unsigned=32) {} };
namespace Style { enum { Default = 7 }; }
struct Event { enum EventType { Closed } type; };
class RenderWindow : public RenderTarget { public: RenderWindow(VideoMode, const std::string&, unsigned=Style::Default, const ContextSettings& = ContextSettings()) {}
  bool isOpen() const { return false; } bool pollEvent(Event&) { return false; } void close() {} void display() {} Vector2u getSize() const override { return {1,1}; } };
struct Keyboard { enum Key { Up, Down, Left, Right }; static bool isKeyPressed(Key) { return false; } };
}  // namespace sf
