// ORACLE — test infrastructure.  An application build in the shape DEMi's README prescribes (README.md:27-29): the
// reference lives in interposition/ and AspectJ weaves akka-actor.
import sbt._
import sbt.Keys._
import com.typesafe.sbt.SbtAspectj.{ Aspectj, aspectjSettings, useInstrumentedClasses }
import com.typesafe.sbt.SbtAspectj.AspectjKeys.inputs

object OracleBuild extends Build {
  lazy val interposition = RootProject(file("interposition"))
  lazy val harness = Project(
    id = "demi-oracle-harness",
    base = file("."),
    settings = Defaults.defaultSettings ++ aspectjSettings ++ Seq(
      scalaVersion := "2.11.2",
      libraryDependencies += "com.typesafe.akka" %% "akka-actor" % "2.3.6",
      inputs in Aspectj <++= update map { report =>
        report.matching(moduleFilter(organization = "com.typesafe.akka", name = "akka-actor*"))
      },
      fullClasspath in Runtime <<= useInstrumentedClasses(Runtime)
    )
  ) dependsOn (interposition)
}
